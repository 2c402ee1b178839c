#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_features_gpu.py tests/test_loop_gpu.py -q -m gpu 2>&1 | tail -4
timeout 600 python bench.py --workload cfg4_stream --steps 20 --warmup 5 --no-stock --no-cpu-baseline --no-alt > gpurun_out/r02k_stream.json 2> gpurun_out/r02k_stream.err
python - <<PY
import json
j = json.load(open("gpurun_out/r02k_stream.json"))
print("stream", {k: v for k, v in j["cfg4_stream"].items() if k not in ("includes", "vs_resident_chunk_note")})
PY
