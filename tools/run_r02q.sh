#!/bin/bash
# (a) selection with 64-bit candidate words + balanced rank counting; (b) weight gradient with 256 x 512 tiles and the A
# slice kept in the tensor core's collector (SCE_TUNE_DW_NSUB2=1, SCE_TUNE_DW_COLL=0/1); (c) wide-d tests
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02q_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02q_pytest.log
SCE_TUNE_DW_NSUB2=1 timeout 900 python -m pytest tests/test_scale_parity_gpu.py -m gpu -x -q -k "config2 or training" > gpurun_out/r02q_pytest_nsub2.log 2>&1; echo "pytest nsub2+coll rc=$?"; tail -3 gpurun_out/r02q_pytest_nsub2.log
show() { python - "$1" "$2" <<'P'
import json,sys
j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[2], round(j["ms_per_step"],4), j.get("phases_ms") or j.get("phases"), j.get("clocks",{}).get("sm_mhz"))
P
}
for rep in 1 2; do
  for v in "0 0" "1 0" "1 1"; do
    set -- $v
    SCE_TUNE_DW_NSUB2=$1 SCE_TUNE_DW_COLL=$2 timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-alt --no-stream --no-stock > gpurun_out/r02q_cfg2_nsub$1_coll$2_$rep.json 2> gpurun_out/r02q_cfg2_nsub$1_coll$2_$rep.err
    show gpurun_out/r02q_cfg2_nsub$1_coll$2_$rep.json "cfg2 nsub2=$1 coll=$2 rep=$rep"
  done
  for w in cfg3 cfg3g; do
    timeout 600 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline --no-alt --no-stream --no-stock > gpurun_out/r02q_${w}_$rep.json 2> gpurun_out/r02q_${w}_$rep.err
    show gpurun_out/r02q_${w}_$rep.json "$w rep=$rep"
  done
done
SCE_TUNE_DW_NSUB2=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"EpiStoreF32" --launch-skip 4 -c 1 -f -o gpurun_out/r02q_dw_coll \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-alt --no-stream --no-stock > gpurun_out/r02q_ncu1.log 2>&1
SCE_TUNE_DW_NSUB2=1 SCE_TUNE_DW_COLL=0 timeout 600 ncu --set full --clock-control none -k regex:"EpiStoreF32" --launch-skip 4 -c 1 -f -o gpurun_out/r02q_dw_nocoll \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-alt --no-stream --no-stock > gpurun_out/r02q_ncu2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"topk_select2" --launch-skip 6 -c 3 -f -o gpurun_out/r02q_select \
  python bench.py --workload cfg3 --steps 1 --warmup 3 --no-cpu-baseline --no-alt --no-stream --no-stock > gpurun_out/r02q_ncu3.log 2>&1
ls -la gpurun_out/r02q_*.ncu-rep
