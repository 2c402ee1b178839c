#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/sanitize_topk.py 2>&1 | tail -4
timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_scale_parity_gpu.py tests/test_features_gpu.py -q -m gpu -k "topk or golden or sweep or determinism or active" 2>&1 | tail -5
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_topk.py 2>&1 | tail -3
for w in cfg3g cfg3; do
  timeout 400 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02g_$w.json 2> gpurun_out/r02g_$w.err
done
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/r02g_*.json")):
    try:
        j = json.load(open(f))
        print(f, "value", round(j["value"]), "ms/step", round(j["ms_per_step"], 3), {k: round(v, 3) for k, v in j.get("phases_ms", {}).items()}, "loss", j["final_loss_mean"])
    except Exception as e:
        print(f, "failed", e); print(open(f[:-5] + ".err").read()[-1500:])
PY
