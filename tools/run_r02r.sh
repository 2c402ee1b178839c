#!/bin/bash
# weight gradient: 256 x 256 tiles (default) against 256 x 512 tiles with the A slice kept in the collector, alternating
mkdir -p gpurun_out
show() { python - "$1" "$2" <<'P'
import json,sys
j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
ph=j.get("phases_ms") or {}
print(sys.argv[2], "step", round(j["ms_per_step"],4), "e2e", round(j["e2e"]["ms_per_step"],4), "dw", round(ph.get("dw",0),4), "sum", round(sum(ph.values()),4), j.get("clocks",{}).get("sm_mhz"))
P
}
for rep in 1 2 3; do
  for v in "0 0" "1 1"; do
    set -- $v
    SCE_TUNE_DW_NSUB2=$1 SCE_TUNE_DW_COLL=$2 timeout 600 python bench.py --steps 60 --warmup 5 --no-cpu-baseline --no-alt --no-stream --no-stock > gpurun_out/r02r_cfg2_nsub$1_coll$2_$rep.json 2> gpurun_out/r02r_cfg2_nsub$1_coll$2_$rep.err
    show gpurun_out/r02r_cfg2_nsub$1_coll$2_$rep.json "cfg2 nsub2=$1 coll=$2 rep=$rep"
  done
done
for v in "1 1" "1 0"; do
  set -- $v
  SCE_TUNE_DW_NSUB2=$1 SCE_TUNE_DW_COLL=$2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_split_kernel --launch-skip 12 -c 4 -f -o gpurun_out/r02r_gemms_nsub$1_coll$2 \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-alt --no-stream --no-stock > gpurun_out/r02r_ncu_$1$2.log 2>&1
done
ls -la gpurun_out/r02r_*.ncu-rep
