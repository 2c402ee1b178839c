#!/bin/bash
# same-box A/B of the two operand arithmetics: parity diagnostics, GPU tests, bench lines
mkdir -p gpurun_out
for a in bf16x3 f16f8; do
  SCE_ARITH=$a timeout 300 python - > gpurun_out/diag_$a.log 2>&1 <<PY
import sys; sys.path.insert(0, "tests")
import diag_parity as D
D.run(2, 128, 256, 1024)
D.run(2, 512, 4096, 2048)
PY
done
cat gpurun_out/diag_bf16x3.log gpurun_out/diag_f16f8.log
timeout 1200 python -m pytest tests/test_engine_gpu.py -q -m gpu -x 2>&1 | tail -30 | tee gpurun_out/f8_pytest.log
SCE_ARITH=bf16x3 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-alt > gpurun_out/ab_bf16x3.json 2> gpurun_out/ab_bf16x3.err
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-alt > gpurun_out/ab_f16f8.json 2> gpurun_out/ab_f16f8.err
python - <<PY
import json
for a in ("bf16x3", "f16f8"):
    try:
        j = json.load(open(f"gpurun_out/ab_{a}.json"))
        print(a, j["config"]["arith"], "ms/step", round(j["ms_per_step"], 3), {k: round(v, 3) for k, v in j["phases_ms"].items()}, "loss", j["final_loss_mean"], j["clocks"])
    except Exception as e:
        print(a, "failed", e); print(open(f"gpurun_out/ab_{a}.err").read()[-2000:])
PY
