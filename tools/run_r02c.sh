#!/bin/bash
# round 2, third GPU pass: k-sparse top-k path — tests, sanitizer, bench A/B against the dense path
mkdir -p gpurun_out
timeout 300 python tools/sanitize_topk.py 2>&1 | tail -5 | tee gpurun_out/r02c_topk_small.log
timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_scale_parity_gpu.py tests/test_features_gpu.py -q -m gpu -k "topk or golden or sweep or determinism or active" 2>&1 | tail -30 | tee gpurun_out/r02c_pytest_topk.log
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_topk.py 2>&1 | tail -15 | tee gpurun_out/r02c_sanitizer_memcheck.log
for sp in 1 0; do
  SCE_TOPK_SPARSE=$sp timeout 400 python bench.py --workload cfg3g --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02c_bench_cfg3g_sparse$sp.json 2> gpurun_out/r02c_bench_cfg3g_sparse$sp.err
done
timeout 400 python bench.py --workload cfg3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02c_bench_cfg3.json 2> gpurun_out/r02c_bench_cfg3.err
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/r02c_bench*.json")):
    try:
        j = json.load(open(f))
        print(f, "value", round(j["value"]), "ms/step", round(j["ms_per_step"], 3), {k: round(v, 3) for k, v in j.get("phases_ms", {}).items()}, "loss", j["final_loss_mean"])
    except Exception as e:
        print(f, "failed", e); print(open(f[:-5] + ".err").read()[-1500:])
PY
