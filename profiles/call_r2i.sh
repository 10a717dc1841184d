#!/bin/bash
# round 2, call i (N GPUs, N = $1): BASELINE configs[3] (c4, 16 M cells, strong scaling) with the per-rank / per-phase wait histogram
N=${1:-8}; shift
OUT=gpurun_out; mkdir -p $OUT
PSM_BENCH_NO_ORACLE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29547 bench.py --gpus $N --steps 30 "$@" > $OUT/bench_r2i_n${N}_c4.json 2> $OUT/bench_r2i_n${N}_c4.err; echo "bench n$N c4 rc=$?"
tail -c 600 $OUT/bench_r2i_n${N}_c4.err
python - <<PY
import json
try:
    b=json.loads(open('gpurun_out/bench_r2i_n${N}_c4.json').read().strip().splitlines()[-1])
    print('n$N c4 ms', round(b['ms_per_step'],4), 'cells/s %.3g' % b['value'], 'e2e', round(b['e2e']['ms_per_step'],4), {k:round(v['ms']*1e3,1) for k,v in b['stages'].items() if v['ms']>0.003}, b.get('parity'), 'init', round(b['init_tables_s'],1), 'limiting', b.get('limiting_phase'))
    for r in b['per_rank']: print(r)
except Exception as e: print('ERR', e)
PY
