#!/bin/bash
# encode / dcode write-out: two adjacent chunks per bulk store (default build) against one chunk per store (build/libsce_nopair.so)
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r02v_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02v_pytest.log
show() { python - "$1" "$2" <<'P'
import json,sys
j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
ph=j.get("phases_ms") or {}
print(sys.argv[2], "step", round(j["ms_per_step"],4), "e2e", round(j["e2e"]["ms_per_step"],4), {k: round(v,3) for k,v in ph.items()}, j.get("clocks",{}).get("sm_mhz"))
P
}
for rep in 1 2 3; do
  for v in pair nopair; do
    if [ $v = nopair ]; then export SCE_LIB=$PWD/build/libsce_nopair.so; else unset SCE_LIB; fi
    timeout 600 python bench.py --steps 60 --warmup 5 --no-cpu-baseline --no-alt --no-stream --no-stock > gpurun_out/r02v_cfg2_${v}_$rep.json 2> gpurun_out/r02v_cfg2_${v}_$rep.err
    show gpurun_out/r02v_cfg2_${v}_$rep.json "cfg2 $v rep=$rep"
  done
done
unset SCE_LIB
for v in pair nopair; do
  if [ $v = nopair ]; then export SCE_LIB=$PWD/build/libsce_nopair.so; else unset SCE_LIB; fi
  timeout 600 python bench.py --workload cfg5 --steps 20 --warmup 5 --no-cpu-baseline --no-alt --no-stream --no-stock > gpurun_out/r02v_cfg5_$v.json 2> gpurun_out/r02v_cfg5_$v.err
  show gpurun_out/r02v_cfg5_$v.json "cfg5 $v"
done
unset SCE_LIB
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_split_kernel --launch-skip 12 -c 4 -f -o gpurun_out/r02v_gemms_pair \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-alt --no-stream --no-stock > gpurun_out/r02v_ncu.log 2>&1
ls -la gpurun_out/r02v_*.ncu-rep
