#!/bin/bash
# weight gradient at d = 768: one double-width (collector) + one single-width tile per row block, interleaved — against three single-width tiles
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r02aa_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02aa_pytest.log
show() { python - "$1" "$2" <<'P'
import json,sys
j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
ph=j.get("phases_ms") or {}
print(sys.argv[2], "step", round(j["ms_per_step"],4), {k: round(v,3) for k,v in ph.items()}, j.get("clocks",{}).get("sm_mhz"))
P
}
for rep in 1 2 3; do
  for v in 0 1; do
    SCE_TUNE_DW_NSUB2=$v timeout 600 python bench.py --workload cfg3 --steps 20 --warmup 5 --no-cpu-baseline --no-alt --no-stream --no-stock > gpurun_out/r02aa_cfg3_nsub${v}_$rep.json 2> gpurun_out/r02aa_cfg3_nsub${v}_$rep.err
    show gpurun_out/r02aa_cfg3_nsub${v}_$rep.json "cfg3 dw_nsub2=$v rep=$rep"
  done
done
