#!/bin/bash
# chunk-maxima selection: tests, A/B on config 3, ncu of the selection kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02p_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02p_pytest.log
for rep in 1 2; do
for cm in 0 1; do
  for w in cfg3 cfg3g; do
    SCE_TOPK_CMAX=$cm timeout 600 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline --no-alt --no-stream --no-stock > gpurun_out/r02p_${w}_cmax${cm}_$rep.json 2> gpurun_out/r02p_${w}_cmax${cm}_$rep.err
    python - <<P
import json
j=json.loads(open("gpurun_out/r02p_${w}_cmax${cm}_$rep.json").read().strip().splitlines()[-1])
print("$w cmax=$cm rep=$rep", j["ms_per_step"], j["value"], j.get("phases"))
P
  done
done
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"topk_select2|EpiScoresTma" --launch-skip 12 -c 6 -f -o gpurun_out/r02p_select \
  python bench.py --workload cfg3 --steps 1 --warmup 3 --no-cpu-baseline --no-alt --no-stream --no-stock > gpurun_out/r02p_ncu.log 2>&1
ls -la gpurun_out/r02p_select.ncu-rep
