#!/bin/bash
# round-end validation on one B200: every GPU test, smoke, both bench arms, the other BASELINE shapes, ncu launch list
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -q -m gpu 2>&1 | tail -6 | tee gpurun_out/final_pytest.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/final_smoke.log
timeout 600 python bench.py --steps 60 --warmup 5 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err
timeout 300 python bench.py --arith bf16x3 --steps 40 --warmup 5 --no-cpu-baseline --no-alt > gpurun_out/final_bench_bf16x3.json 2>/dev/null
timeout 300 python bench.py --act-precision fp32 --steps 40 --warmup 5 --no-cpu-baseline --no-alt > gpurun_out/final_bench_fp32vals.json 2>/dev/null
for w in cfg1 cfg3 cfg5; do
  timeout 400 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline --no-alt > gpurun_out/final_bench_$w.json 2> gpurun_out/final_bench_$w.err
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-alt > gpurun_out/final_ncu_bench.log 2>&1
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/final_bench*.json")):
    try:
        j = json.load(open(f))
        print(f[11:-5].ljust(22), j.get("impl", j["config"].get("arith")), "value", round(j["value"]), "ms/step", round(j["ms_per_step"], 3),
              "e2e", round(j["e2e"]["value"]), {k: round(v, 3) for k, v in j.get("phases_ms", {}).items()})
    except Exception as e:
        print(f, "failed", e)
PY
