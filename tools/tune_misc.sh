#!/bin/bash
run() {
  env "$@" timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); p=d['phases_ms']
print('%-50s step %.3f ms  enc %.3f dec %.3f dcode %.3f dw %.3f  sm %s MHz' % ('$*', d['ms_per_step'], p['encode'], p['decode'], p['dcode'], p['dw'], d['clocks']['sm_mhz']))"
}
run SCE_X=0
run SCE_TUNE_SPLIT_DECODE=0
run SCE_TUNE_SPLIT_DECODE=0 SCE_TUNE_BK_DECODE=64
run SCE_X=0
run SCE_TUNE_SPLIT_DECODE=0
