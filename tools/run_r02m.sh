#!/bin/bash
mkdir -p gpurun_out
SCE_TOPK_SPARSE=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:topk_sparse --launch-skip 4 -c 1 -f -o gpurun_out/r02m_sparse \
  python bench.py --workload cfg3g --steps 1 --warmup 3 --no-cpu-baseline --no-alt --no-stream --no-stock > gpurun_out/r02m_ncu.log 2>&1
ls -la gpurun_out/r02m_sparse.ncu-rep
