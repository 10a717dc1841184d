#!/bin/bash
# round 2, call o (2 GPUs): sharded parity after the send-map refactor, then the weak workload (c2 x 2) with the wait histogram
OUT=gpurun_out; mkdir -p $OUT
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/mgpu_worker.py --variant deltaU_to_deltaP --builder band > $OUT/mgpu_r2o_n2.log 2>&1; echo "mgpu world 2 rc=$?"; grep -E "mgpu |routed|MGPU" $OUT/mgpu_r2o_n2.log | cut -c1-300
PSM_BENCH_NO_ORACLE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29547 bench.py --gpus 2 --workload c2xN --steps 20 --no-parity > $OUT/bench_r2o_n2.json 2> $OUT/bench_r2o_n2.err; echo "bench n2 rc=$?"
python - <<PY
import json
try:
    b=json.loads(open('gpurun_out/bench_r2o_n2.json').read().strip().splitlines()[-1])
    print('n2 c2x2 ms', round(b['ms_per_step'],4), 'cells/s %.3g' % b['value'], 'e2e', round(b['e2e']['ms_per_step'],4), {k:round(v['ms']*1e3,1) for k,v in b['stages'].items() if v['ms']>0.003})
    for r in b['per_rank']: print(r['rank'], r['cells'], round(r['ms_per_step'],4), 'noflush', round(r['ms_per_step_no_flush'],4), {k:round(v,1) for k,v in r['wait_us_per_step'].items()})
except Exception as e: print('ERR', e)
PY
