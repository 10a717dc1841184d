#!/bin/bash
# round 2, call l (N GPUs): ghost-cell send map -- parity of the sharded flow (world N, band builder, cell routing), then c4 strong
N=${1:-4}
OUT=gpurun_out; mkdir -p $OUT
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tests/mgpu_worker.py --variant deltaU_to_deltaP --builder band > $OUT/mgpu_r2l_n$N.log 2>&1; echo "mgpu world $N rc=$?"; grep -E "mgpu |routed|MGPU" $OUT/mgpu_r2l_n$N.log | cut -c1-400
PSM_BENCH_NO_ORACLE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29547 bench.py --gpus $N --steps 30 --no-parity > $OUT/bench_r2l_n${N}_c4.json 2> $OUT/bench_r2l_n${N}_c4.err; echo "bench n$N c4 rc=$?"
tail -c 400 $OUT/bench_r2l_n${N}_c4.err
python - <<PY
import json
try:
    b=json.loads(open('gpurun_out/bench_r2l_n${N}_c4.json').read().strip().splitlines()[-1])
    print('n$N c4 ms', round(b['ms_per_step'],4), 'cells/s %.3g' % b['value'], 'e2e', round(b['e2e']['ms_per_step'],4), {k:round(v['ms']*1e3,1) for k,v in b['stages'].items() if v['ms']>0.003}, 'init', round(b['init_tables_s'],1), 'limiting', b.get('limiting_phase'))
    for r in b['per_rank']: print(r['rank'], r['cells'], round(r['ms_per_step'],4), 'noflush', round(r['ms_per_step_no_flush'],4), {k:round(v,1) for k,v in r['wait_us_per_step'].items()})
except Exception as e: print('ERR', e)
PY
