#!/bin/bash
# quick GPU check of the current build: the GPU parity tests, then one bench line per arithmetic given in $1..
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_engine_gpu.py -q -m gpu -x 2>&1 | tail -15 | tee gpurun_out/quick_pytest.log
for a in "$@"; do
  timeout 300 python bench.py --arith $a --steps 40 --warmup 5 --no-cpu-baseline --no-alt > gpurun_out/quick_$a.json 2> gpurun_out/quick_$a.err
done
python - "$@" <<PY
import json, sys
for a in sys.argv[1:]:
    try:
        j = json.load(open(f"gpurun_out/quick_{a}.json"))
        print(a, j["config"]["arith"], "ms/step", round(j["ms_per_step"], 3), "e2e", round(j["e2e"]["ms_per_step"], 3), {k: round(v, 3) for k, v in j["phases_ms"].items()}, "loss", j["final_loss_mean"], j["clocks"])
    except Exception as e:
        print(a, "failed", e); print(open(f"gpurun_out/quick_{a}.err").read()[-2000:])
PY
