#!/bin/bash
# 2 GPUs: streamed config 4 with both feeds (per-rank PCIe copies vs one H2D + NCCL broadcast), then the default line
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 2 --workload cfg4_stream --feed both --steps 20 --warmup 5 --no-stock --no-cpu-baseline --no-alt \
  > gpurun_out/r02e_stream_2gpu.json 2> gpurun_out/r02e_stream_2gpu.err
tail -3 gpurun_out/r02e_stream_2gpu.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02e_bench_2gpu.json 2> gpurun_out/r02e_bench_2gpu.err
tail -3 gpurun_out/r02e_bench_2gpu.err
python - <<PY
import json
for f in ("gpurun_out/r02e_stream_2gpu.json", "gpurun_out/r02e_bench_2gpu.json"):
    try:
        j = json.load(open(f))
        print(f, "value", round(j["value"]), "ms/step", round(j["ms_per_step"], 3), "e2e", round(j["e2e"]["value"]))
        cs = j.get("cfg4_stream")
        print("   stream:", json.dumps(cs)[:1800])
    except Exception as e:
        print(f, "failed", e)
PY
