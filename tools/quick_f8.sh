#!/bin/bash
# quick GPU check of the current build: GEMM self-test (f16f8 cases), the GPU parity tests, then one bench line per
# "arith[:act-precision]" argument
mkdir -p gpurun_out
timeout 200 ./build/gemm_selftest --f8 2>&1 | tee gpurun_out/quick_selftest.log | grep -v PASS
timeout 1200 python -m pytest tests/test_engine_gpu.py -q -m gpu -x 2>&1 | tail -15 | tee gpurun_out/quick_pytest.log
for a in "$@"; do
  ar=${a%%:*}; ap=${a##*:}; [ "$ap" = "$a" ] && ap=fp16
  timeout 300 python bench.py --arith $ar --act-precision $ap --steps 40 --warmup 5 --no-cpu-baseline --no-alt > gpurun_out/quick_${ar}_$ap.json 2> gpurun_out/quick_${ar}_$ap.err
done
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/quick_*_fp*.json")):
    try:
        j = json.load(open(f))
        print(f[17:-5], j["config"]["arith"], "ms/step", round(j["ms_per_step"], 3), "e2e", round(j["e2e"]["ms_per_step"], 3), {k: round(v, 3) for k, v in j["phases_ms"].items()}, "loss", j["final_loss_mean"], j["clocks"]["sm_mhz"])
    except Exception as e:
        print(f, "failed", e); print(open(f[:-5] + ".err").read()[-2000:])
PY
