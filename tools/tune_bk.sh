#!/bin/bash
# A/B the K block of the three K-major GEMMs inside ONE gpurun call (same box, same thermal state).
# usage: tools/tune_bk.sh > gpurun_out/bk_tuning.txt
for cfg in "64 64 64" "64 32 64" "32 32 32" "32 32 64" "64 32 32" "64 64 64" "64 32 64"; do
  set -- $cfg
  SCE_TUNE_BK_ENCODE=$1 SCE_TUNE_BK_DECODE=$2 SCE_TUNE_BK_DCODE=$3 python bench.py --steps 40 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); p=d['phases_ms']
print('enc/dec/dcode BK = $1/$2/$3  step %.3f ms  encode %.3f decode %.3f dcode %.3f dw %.3f  sm %s MHz' % (d['ms_per_step'], p['encode'], p['decode'], p['dcode'], p['dw'], d['clocks']['sm_mhz']))"
done
