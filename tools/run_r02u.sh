#!/bin/bash
# 8 GPUs: the default bench line exactly as the driver launches it (with the config-4 streaming extras)
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 \
  bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02u_bench_8gpu.json 2> gpurun_out/r02u_bench_8gpu.err
tail -3 gpurun_out/r02u_bench_8gpu.err
python - <<PY
import json
j = json.load(open("gpurun_out/r02u_bench_8gpu.json"))
print("value", round(j["value"]), "ms/step", round(j["ms_per_step"], 3), "e2e", round(j["e2e"]["value"]), j["clocks"])
print("stream:", json.dumps({k: v for k, v in j.get("cfg4_stream", {}).items() if k not in ("includes", "vs_resident_chunk_note")}))
PY
