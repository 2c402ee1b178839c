#!/bin/bash
# experiment: decode with 256 x 512 tiles + collector (SCE_TUNE_DEC_NSUB2=1)
mkdir -p gpurun_out
SCE_TUNE_DEC_NSUB2=1 timeout 900 python -m pytest tests/test_scale_parity_gpu.py tests/test_engine_gpu.py -m gpu -x -q -k "config2 or golden or trajectory" > gpurun_out/r02y_pytest.log 2>&1; echo "pytest (dec nsub2) rc=$?"; tail -3 gpurun_out/r02y_pytest.log
show() { python - "$1" "$2" <<'P'
import json,sys
j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
ph=j.get("phases_ms") or {}
print(sys.argv[2], "step", round(j["ms_per_step"],4), "e2e", round(j["e2e"]["ms_per_step"],4), {k: round(v,3) for k,v in ph.items()}, j.get("clocks",{}).get("sm_mhz"))
P
}
for rep in 1 2 3; do
  for v in 0 1; do
    SCE_TUNE_DEC_NSUB2=$v timeout 600 python bench.py --steps 60 --warmup 5 --no-cpu-baseline --no-alt --no-stream --no-stock > gpurun_out/r02y_cfg2_dec${v}_$rep.json 2> gpurun_out/r02y_cfg2_dec${v}_$rep.err
    show gpurun_out/r02y_cfg2_dec${v}_$rep.json "cfg2 dec_nsub2=$v rep=$rep"
  done
done
