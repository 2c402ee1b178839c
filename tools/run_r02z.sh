#!/bin/bash
# last check of the committed tree: GPU suite, smoke, default bench line
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r02z_pytest.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/r02z_bench.json 2> gpurun_out/r02z_bench.err; python - <<'P'
import json
j=json.loads(open("gpurun_out/r02z_bench.json").read().strip().splitlines()[-1])
print("value", round(j["value"]), "ms", round(j["ms_per_step"],3), "e2e", round(j["e2e"]["value"]), "frac", round(j["roofline"]["frac"],3), "traffic", j["roofline"]["traffic"], "launches", j["gpu_launches"], j["clocks"]["sm_mhz"], j["clocks"]["reasons"])
P
