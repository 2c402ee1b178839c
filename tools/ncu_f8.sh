#!/bin/bash
# one ncu --set full capture of the four GEMMs (+ the streaming kernels between them) of one warm step
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on --launch-skip 30 -c 12 -f -o gpurun_out/$1 \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-alt ${@:2} > gpurun_out/$1.log 2>&1
ls -la gpurun_out/$1.ncu-rep
