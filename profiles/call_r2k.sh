#!/bin/bash
# round 2, call k (one GPU): grid-plane A operand with warp-wide box issue, Dense epilogue vectors prefetched, one-wave Dense sizing
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r2k}
timeout 240 python -m pytest tests/test_gpu_parity.py -x -q > $OUT/pytest_${TAG}_quick.log 2>&1; echo "quick rc=$?"; tail -2 $OUT/pytest_${TAG}_quick.log
PSM_GRID_A=1 timeout 240 python -m pytest tests/test_gpu_parity.py -x -q > $OUT/pytest_${TAG}_grid_quick.log 2>&1; rc=$?; echo "grid quick rc=$rc"; tail -2 $OUT/pytest_${TAG}_grid_quick.log
timeout 300 python bench.py --no-cpu-baseline --steps 30 > $OUT/ab_${TAG}_default_c2.json 2>> $OUT/ab_$TAG.err; echo "default c2 rc=$?"
timeout 300 python bench.py --variant U_to_gradP --no-cpu-baseline --no-c5 --steps 30 > $OUT/ab_${TAG}_default_c3.json 2>> $OUT/ab_$TAG.err; echo "default c3 rc=$?"
if [ $rc -eq 0 ]; then
PSM_GRID_A=1 timeout 300 python bench.py --no-cpu-baseline --steps 30 > $OUT/ab_${TAG}_grid_c2.json 2>> $OUT/ab_$TAG.err; echo "grid c2 rc=$?"
PSM_GRID_A=1 timeout 300 python bench.py --variant U_to_gradP --no-cpu-baseline --no-c5 --steps 30 > $OUT/ab_${TAG}_grid_c3.json 2>> $OUT/ab_$TAG.err; echo "grid c3 rc=$?"
fi
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/ab_${TAG}_*.json')):
    try:
        b=json.loads(open(f).read().strip().splitlines()[-1])
        e=b['e2e']; c5=b.get('roofline_c5') or {}
        print(f.split('/')[-1], 'ms', round(b['ms_per_step'],4), 'e2e', round(e['ms_per_step'],4), {k:round(v['ms']*1e3,1) for k,v in b['stages'].items() if v['ms']>0.003}, 'c5', c5.get('ms_per_step'), {k:round(v*1e3,1) for k,v in (c5.get('stages_ms') or {}).items() if v>0.003})
    except Exception as ex: print(f, 'ERR', ex)
PY
tail -5 $OUT/ab_$TAG.err
