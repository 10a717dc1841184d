#!/bin/bash
# round 2, call c (2 GPUs): 2-rank parity tests, bench c2 / c3, the 2-GPU weak bench and c4 (16 M cells) over 2 GPUs
OUT=gpurun_out; mkdir -p $OUT
python -m pytest tests/test_gpu_multi.py -x -q -k two_rank > $OUT/pytest_r2c_multi.log 2>&1; echo "multi rc=$?"; tail -12 $OUT/pytest_r2c_multi.log | cut -c1-300
python -m pytest tests/test_gpu_parity.py tests/test_gpu_files.py -x -q > $OUT/pytest_r2c_single.log 2>&1; echo "single rc=$?"; tail -4 $OUT/pytest_r2c_single.log
python bench.py --no-c5 --steps 30 > $OUT/bench_r2c_c2.json 2> $OUT/bench_r2c_c2.err; echo "bench c2 rc=$?"
python bench.py --no-c5 --steps 20 --variant U_to_gradP --no-cpu-baseline > $OUT/bench_r2c_c3.json 2> $OUT/bench_r2c_c3.err; echo "bench c3 rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 --workload c2xN --steps 30 > $OUT/bench_r2c_n2.json 2> $OUT/bench_r2c_n2.err; echo "bench n2 weak rc=$?"
PSM_BENCH_NO_ORACLE=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29545 bench.py --gpus 2 --steps 20 > $OUT/bench_r2c_n2_c4.json 2> $OUT/bench_r2c_n2_c4.err; echo "bench n2 c4 rc=$?"
tail -c 400 $OUT/bench_r2c_n2_c4.err
python - <<'PY'
import json
for f in ['bench_r2c_c2','bench_r2c_c3','bench_r2c_n2','bench_r2c_n2_c4']:
    try:
        b=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
        print(f, 'ms', round(b['ms_per_step'],4), 'cells/s %.3g' % b['value'], 'e2e', round(b['e2e']['ms_per_step'],4), {k:round(v['ms']*1e3,1) for k,v in b['stages'].items()}, b.get('parity'), 'init', round(b['init_tables_s'],1))
    except Exception as e: print(f, 'ERR', e)
PY
