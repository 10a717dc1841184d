#!/bin/bash
# round 2, call n (one GPU): projection split-K rounded down to one resident wave (c3: 7 tiles x 21 splits = 147 CTAs, was 154)
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r2n}
timeout 200 python -m pytest tests/test_gpu_parity.py -x -q > $OUT/pytest_${TAG}_quick.log 2>&1; echo "quick rc=$?"; tail -2 $OUT/pytest_${TAG}_quick.log
timeout 200 python bench.py --variant U_to_gradP --no-cpu-baseline --no-c5 --steps 30 > $OUT/bench_${TAG}_c3.json 2> $OUT/bench_${TAG}_c3.err; echo "bench c3 rc=$?"
timeout 200 python bench.py --no-cpu-baseline --no-c5 --steps 30 > $OUT/bench_${TAG}_c2.json 2> $OUT/bench_${TAG}_c2.err; echo "bench c2 rc=$?"
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_${TAG}_*.json')):
    try:
        b=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], 'ms', round(b['ms_per_step'],4), 'e2e', round(b['e2e']['ms_per_step'],4), {k:round(v['ms']*1e3,1) for k,v in b['stages'].items() if v['ms']>0.003})
    except Exception as ex: print(f, 'ERR', ex)
PY
