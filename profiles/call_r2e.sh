#!/bin/bash
# round 2, call e (2 GPUs): 2-rank tests incl. cell routing, single-GPU parity incl. the forced strip fusion, 2-GPU weak bench
OUT=gpurun_out; mkdir -p $OUT
python -m pytest tests/test_gpu_multi.py -x -q -k two_rank > $OUT/pytest_r2e_multi.log 2>&1; echo "multi rc=$?"; tail -14 $OUT/pytest_r2e_multi.log | cut -c1-400
python -m pytest tests/test_gpu_parity.py tests/test_gpu_files.py -x -q > $OUT/pytest_r2e_single.log 2>&1; echo "single rc=$?"; tail -6 $OUT/pytest_r2e_single.log | cut -c1-400
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 --workload c2xN --steps 30 > $OUT/bench_r2e_n2.json 2> $OUT/bench_r2e_n2.err; echo "bench n2 weak rc=$?"
python bench.py --no-c5 --steps 30 --no-cpu-baseline > $OUT/bench_r2e_c2.json 2> $OUT/bench_r2e_c2.err; echo "bench c2 rc=$?"
python - <<'PY'
import json
for f in ['bench_r2e_c2','bench_r2e_n2']:
    try:
        b=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
        print(f, 'ms', round(b['ms_per_step'],4), 'cells/s %.3g' % b['value'], 'e2e', round(b['e2e']['ms_per_step'],4), {k:round(v['ms']*1e3,1) for k,v in b['stages'].items() if v['ms']>0.003}, b.get('parity'))
    except Exception as e: print(f, 'ERR', e)
PY
