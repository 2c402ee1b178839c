#!/bin/bash
# round 2 FINAL validation pass (device-side centring, chunk-maxima selection, collector dW tiles): every GPU test, smoke, bench lines of every workload, both arms, ncu launch list + full capture
mkdir -p gpurun_out
rm -f gpurun_out/r02_parity_report.txt
timeout 1500 python -m pytest tests/ -q -m gpu 2>&1 | tail -8 | tee gpurun_out/r02x_pytest.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/r02x_smoke.log
( time timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02x_bench.json 2> gpurun_out/r02x_bench.err ) 2> gpurun_out/r02x_bench_time.log; cat gpurun_out/r02x_bench_time.log
timeout 600 python bench.py --steps 60 --warmup 5 --no-stream --no-stock --no-cpu-baseline > gpurun_out/r02x_bench_60steps.json 2> gpurun_out/r02x_bench_60steps.err
timeout 400 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02x_bench_ref.json 2> gpurun_out/r02x_bench_ref.err
timeout 300 python bench.py --arith bf16x3 --steps 40 --warmup 5 --no-cpu-baseline --no-alt --no-stream --no-stock > gpurun_out/r02x_bench_bf16x3.json 2>/dev/null
timeout 300 python bench.py --act-precision fp32 --steps 40 --warmup 5 --no-cpu-baseline --no-alt --no-stream --no-stock > gpurun_out/r02x_bench_fp32vals.json 2>/dev/null
for w in cfg1 cfg5 cfg3 cfg3g; do
  timeout 400 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline --no-alt --no-stock > gpurun_out/r02x_bench_$w.json 2> gpurun_out/r02x_bench_$w.err
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02x_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-alt --no-stream --no-stock > gpurun_out/r02x_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --launch-skip 44 -c 12 -f -o gpurun_out/r02x_full \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-alt --no-stream --no-stock > gpurun_out/r02x_ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:topk --launch-skip 6 -c 3 -f -o gpurun_out/r02x_topk_full \
  python bench.py --workload cfg3 --steps 1 --warmup 3 --no-cpu-baseline --no-alt --no-stream --no-stock > gpurun_out/r02x_ncu_topk.log 2>&1
timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_topk.py > gpurun_out/r02x_racecheck_topk.log 2>&1; tail -3 gpurun_out/r02x_racecheck_topk.log
ls -la gpurun_out/*.ncu-rep | tail -3
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/r02x_bench*.json")):
    try:
        j = json.load(open(f))
        print(f[11:-5].ljust(26), j.get("impl", j["config"].get("arith")), "value", round(j["value"]), "ms/step", round(j["ms_per_step"], 3),
              "e2e", round(j["e2e"]["value"]), {k: round(v, 3) for k, v in j.get("phases_ms", {}).items()})
    except Exception as e:
        print(f, "failed", e)
PY
