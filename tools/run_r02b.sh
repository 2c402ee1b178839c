#!/bin/bash
# round 2, second GPU pass: epilogue write-out variants (same-box A/B), tests on the LSU variant, streamed config 4
mkdir -p gpurun_out
BENCH_ARGS="--no-stock --no-stream" TEST_VARIANT=lsu bash tools/ab_variants.sh 2>&1 | tee gpurun_out/r02b_variants.txt
BENCH_ARGS="--no-stock --no-stream" bash tools/ab_variants.sh 2>&1 | tee -a gpurun_out/r02b_variants.txt
timeout 600 python bench.py --workload cfg4_stream --steps 20 --warmup 5 --no-stock --no-cpu-baseline --no-alt > gpurun_out/r02b_stream.json 2> gpurun_out/r02b_stream.err
python - <<PY
import json
j = json.load(open("gpurun_out/r02b_stream.json"))
print("stream", {k: v for k, v in j["cfg4_stream"].items() if k not in ("includes", "vs_resident_chunk_note")})
PY
tail -5 gpurun_out/r02b_stream.err
