#!/bin/bash
# round 2, first GPU pass: every GPU test (incl. the config-scale parity file), smoke, bench lines
mkdir -p gpurun_out
rm -f gpurun_out/r02_parity_report.txt
df -h /dev/shm /tmp | tee gpurun_out/r02a_env.txt; nproc >> gpurun_out/r02a_env.txt; free -g >> gpurun_out/r02a_env.txt; nvidia-smi -L >> gpurun_out/r02a_env.txt
timeout 1500 python -m pytest tests/ -q -m gpu 2>&1 | tail -40 | tee gpurun_out/r02a_pytest.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/r02a_smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err
tail -3 gpurun_out/r02a_bench.err
timeout 400 python bench.py --workload cfg3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02a_bench_cfg3.json 2> gpurun_out/r02a_bench_cfg3.err
timeout 400 python bench.py --workload cfg3g --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02a_bench_cfg3g.json 2> gpurun_out/r02a_bench_cfg3g.err
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/r02a_bench*.json")):
    try:
        j = json.load(open(f))
        print(f, "value", round(j["value"]), "ms/step", round(j["ms_per_step"], 3), "e2e", round(j["e2e"]["value"]), j["e2e"].get("serial"),
              {k: round(v, 3) for k, v in j.get("phases_ms", {}).items()}, j.get("clocks"))
        for k in ("cfg4_stream", "stock_torch_gpu", "cpu_baseline", "alt_precision"):
            print("   ", k, j.get(k))
    except Exception as e:
        print(f, "failed", e); print(open(f[:-5] + ".err").read()[-1500:])
PY
