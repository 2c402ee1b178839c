#!/bin/bash
# gather kernel launched per k class (shared memory sized by the class, not by k_max): tests, memcheck, config 3 lines
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r02w_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02w_pytest.log
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_topk.py > gpurun_out/r02w_memcheck_topk.log 2>&1; tail -3 gpurun_out/r02w_memcheck_topk.log
show() { python - "$1" "$2" <<'P'
import json,sys
j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
ph=j.get("phases_ms") or {}
print(sys.argv[2], "step", round(j["ms_per_step"],4), {k: round(v,3) for k,v in ph.items()}, j.get("clocks",{}).get("sm_mhz"))
P
}
for rep in 1 2; do
  for w in cfg3 cfg3g; do
    timeout 600 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline --no-alt --no-stream --no-stock > gpurun_out/r02w_${w}_$rep.json 2> gpurun_out/r02w_${w}_$rep.err
    show gpurun_out/r02w_${w}_$rep.json "$w rep=$rep"
  done
  SCE_TOPK_SPARSE=1 timeout 600 python bench.py --workload cfg3g --steps 20 --warmup 5 --no-cpu-baseline --no-alt --no-stream --no-stock > gpurun_out/r02w_cfg3g_sparse_$rep.json 2> gpurun_out/r02w_cfg3g_sparse_$rep.err
  show gpurun_out/r02w_cfg3g_sparse_$rep.json "cfg3g forced sparse rep=$rep"
  SCE_TOPK_SPARSE=1 timeout 600 python bench.py --workload cfg3 --steps 20 --warmup 5 --no-cpu-baseline --no-alt --no-stream --no-stock > gpurun_out/r02w_cfg3_sparse_$rep.json 2> gpurun_out/r02w_cfg3_sparse_$rep.err
  show gpurun_out/r02w_cfg3_sparse_$rep.json "cfg3 forced sparse (all groups) rep=$rep"
done
