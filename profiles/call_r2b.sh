#!/bin/bash
# round 2, call b (2 GPUs): changed single-GPU tests, the 2-rank parity tests, bench c2 and the 2-GPU weak bench
OUT=gpurun_out; mkdir -p $OUT
python -m pytest tests/test_gpu_parity.py tests/test_gpu_files.py tests/test_gpu_thesis.py tests/test_gpu_gemm.py -x -q > $OUT/pytest_r2b_single.log 2>&1; echo "single rc=$?"; tail -4 $OUT/pytest_r2b_single.log
python -m pytest tests/test_gpu_multi.py -x -q -k two_rank > $OUT/pytest_r2b_multi.log 2>&1; echo "multi rc=$?"; tail -30 $OUT/pytest_r2b_multi.log | cut -c1-400
python bench.py --no-c5 --steps 30 > $OUT/bench_r2b_c2.json 2> $OUT/bench_r2b_c2.err; echo "bench c2 rc=$?"
python bench.py --no-c5 --steps 20 --variant U_to_gradP --no-cpu-baseline > $OUT/bench_r2b_c3.json 2> $OUT/bench_r2b_c3.err; echo "bench c3 rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 --workload c2xN --steps 30 > $OUT/bench_r2b_n2.json 2> $OUT/bench_r2b_n2.err; echo "bench n2 rc=$?"
tail -c 600 $OUT/bench_r2b_n2.err
python - <<'PY'
import json
for f in ['bench_r2b_c2','bench_r2b_c3','bench_r2b_n2']:
    try:
        b=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
        print(f, 'ms', round(b['ms_per_step'],4), 'e2e', round(b['e2e']['ms_per_step'],4), {k:round(v['ms']*1e3,1) for k,v in b['stages'].items()}, b.get('parity'))
    except Exception as e: print(f, 'ERR', e)
PY
