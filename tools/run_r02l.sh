#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/sanitize_topk.py 2>&1 | tail -4
SCE_TOPK_SPARSE=1 timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_scale_parity_gpu.py -q -m gpu -k "topk or golden or sweep or determinism" 2>&1 | tail -3
SCE_TOPK_SPARSE=1 timeout 400 python bench.py --workload cfg3g --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02l_cfg3g_sparse.json 2> gpurun_out/r02l_cfg3g_sparse.err
SCE_TOPK_SPARSE=0 timeout 400 python bench.py --workload cfg3g --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02l_cfg3g_dense.json 2> gpurun_out/r02l_cfg3g_dense.err
SCE_TOPK_SPARSE=1 timeout 400 python bench.py --workload cfg3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02l_cfg3_allsparse.json 2> gpurun_out/r02l_cfg3_allsparse.err
timeout 400 python bench.py --workload cfg3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02l_cfg3_heuristic.json 2> gpurun_out/r02l_cfg3_heuristic.err
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/r02l_*.json")):
    try:
        j = json.load(open(f))
        print(f, "value", round(j["value"]), "ms/step", round(j["ms_per_step"], 3), {k: round(v, 3) for k, v in j.get("phases_ms", {}).items()}, "loss", j["final_loss_mean"])
    except Exception as e:
        print(f, "failed", e); print(open(f[:-5] + ".err").read()[-1500:])
PY
