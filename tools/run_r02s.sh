#!/bin/bash
# device-side centring + double-width dW tiles with collector by default: full GPU suite, config 2 / 3 / 5 lines
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r02s_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02s_pytest.log
show() { python - "$1" "$2" <<'P'
import json,sys
j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
ph=j.get("phases_ms") or {}
print(sys.argv[2], "step", round(j["ms_per_step"],4), "e2e", round(j["e2e"]["ms_per_step"],4), {k: round(v,3) for k,v in ph.items()}, j.get("clocks",{}).get("sm_mhz"))
P
}
for w in cfg2 cfg5 cfg3 cfg3g; do
  if [ $w = cfg2 ]; then a=""; else a="--workload $w"; fi
  timeout 600 python bench.py $a --steps 20 --warmup 5 --no-cpu-baseline --no-alt --no-stream --no-stock > gpurun_out/r02s_$w.json 2> gpurun_out/r02s_$w.err
  show gpurun_out/r02s_$w.json $w
done
SCE_TUNE_DW_NSUB2=0 timeout 600 python bench.py --workload cfg5 --steps 20 --warmup 5 --no-cpu-baseline --no-alt --no-stream --no-stock > gpurun_out/r02s_cfg5_nsub0.json 2> gpurun_out/r02s_cfg5_nsub0.err
show gpurun_out/r02s_cfg5_nsub0.json "cfg5 nsub2=0"
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_engine_gpu.py -m gpu -x -q -k "centring or wide_activation or other_config_shapes" > gpurun_out/r02s_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -5 gpurun_out/r02s_memcheck.log
