#!/bin/bash
# round 2, call f (N GPUs, N = $1): one world-N parity test (band builder, fused flow, cell routing), then BASELINE configs[3]
# (c4, 16 M cells) over N GPUs with the parity check against a single-GPU handle inside the run
N=${1:-8}
OUT=gpurun_out; mkdir -p $OUT
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tests/mgpu_worker.py --variant deltaU_to_deltaP --builder band > $OUT/mgpu_r2f_n$N.log 2>&1; echo "mgpu world $N rc=$?"; grep -E "mgpu |routed|MGPU" $OUT/mgpu_r2f_n$N.log | cut -c1-400
PSM_BENCH_NO_ORACLE=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29545 bench.py --gpus $N --steps 30 > $OUT/bench_r2f_n${N}_c4.json 2> $OUT/bench_r2f_n${N}_c4.err; echo "bench n$N c4 rc=$?"
tail -c 300 $OUT/bench_r2f_n${N}_c4.err
python - <<PY
import json
try:
    b=json.loads(open('gpurun_out/bench_r2f_n${N}_c4.json').read().strip().splitlines()[-1])
    print('n$N c4 ms', round(b['ms_per_step'],4), 'cells/s %.3g' % b['value'], 'e2e', round(b['e2e']['ms_per_step'],4), {k:round(v['ms']*1e3,1) for k,v in b['stages'].items() if v['ms']>0.003}, b.get('parity'), 'init', round(b['init_tables_s'],1), b['transport'])
except Exception as e: print('ERR', e)
PY
