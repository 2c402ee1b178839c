#!/bin/bash
# same-box A/B of libsce.so builds under build/var/ (see tools/build_variants.sh): bench phases per variant
mkdir -p gpurun_out
for so in build/var/libsce_*.so; do
  tag=$(basename $so .so); tag=${tag#libsce_}
  SCE_LIB=$PWD/$so timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-alt ${BENCH_ARGS} > gpurun_out/var_$tag.json 2> gpurun_out/var_$tag.err
done
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/var_*.json")):
    try:
        j = json.load(open(f))
        print(f[15:-5].ljust(20), j["config"]["arith"], "ms/step", round(j["ms_per_step"], 3), {k: round(v, 3) for k, v in j["phases_ms"].items()}, "loss", round(j["final_loss_mean"], 6))
    except Exception as e:
        print(f, "failed", e); print(open(f[:-5] + ".err").read()[-1500:])
PY
if [ -n "$TEST_VARIANT" ]; then
  SCE_LIB=$PWD/build/var/libsce_$TEST_VARIANT.so timeout 900 python -m pytest tests/test_engine_gpu.py -q -m gpu -x 2>&1 | tail -8
fi
