#!/bin/bash
# store-warp epilogue: correctness under a short timeout first, then same-box A/B against the previous build
mkdir -p gpurun_out
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 600 python -m pytest tests/test_engine_gpu.py -q -m gpu -x 2>&1 | tail -5
timeout 600 python -m pytest tests/test_scale_parity_gpu.py tests/test_features_gpu.py tests/test_loop_gpu.py -q -m gpu 2>&1 | tail -5
BENCH_ARGS="--no-stock --no-stream" bash tools/ab_variants.sh 2>&1 | tee gpurun_out/r02i_variants.txt
BENCH_ARGS="--no-stock --no-stream" bash tools/ab_variants.sh 2>&1 | tee -a gpurun_out/r02i_variants.txt
