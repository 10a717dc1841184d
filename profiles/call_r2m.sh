#!/bin/bash
# round 2, call m (one GPU): the round's final capture.  GPU parity tests -> default bench line (c2 + roofline_c5 + cpu_baseline)
# -> configs[2] (c3) -> configs[0] (c1) -> configs[4] (1000-step latency run on c5) -> ncu launch list -> ncu --set full of one step.
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r2m}
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_$TAG.log
timeout 600 python bench.py > $OUT/bench_${TAG}_c2.json 2> $OUT/bench_${TAG}_c2.err; echo "bench c2 rc=$?"
timeout 300 python bench.py --variant U_to_gradP --no-cpu-baseline --no-c5 > $OUT/bench_${TAG}_c3.json 2> $OUT/bench_${TAG}_c3.err; echo "bench c3 rc=$?"
timeout 300 python bench.py --workload c1 --no-cpu-baseline --no-c5 --steps 30 > $OUT/bench_${TAG}_c1.json 2> $OUT/bench_${TAG}_c1.err; echo "bench c1 rc=$?"
timeout 300 python bench.py --workload c5 --latency --steps 1000 --no-cpu-baseline > $OUT/bench_${TAG}_c5_latency.json 2> $OUT/bench_${TAG}_c5_latency.err; echo "latency rc=$?"
K='regex:^(void )?(psm::)?(prep|gather|extract|tc_gemm|proj_|dense_cluster|dense_stack|pca_inverse_t|reduce_standardise|task_|offsets|place|back|gauss|p2p|init_)'
CMD="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-c5"
$CMD > $OUT/plain_$TAG.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches_$TAG.csv $CMD > $OUT/ncu_launch_$TAG.log 2>&1
echo "launch list rc=$?"
$CMD > $OUT/plain2_$TAG.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k "$K" -s ${NCU_SKIP:-44} -c ${NCU_COUNT:-11} -f -o $OUT/prof_$TAG $CMD > $OUT/ncu_full_$TAG.log 2>&1
echo "full capture rc=$?"
CMD3="python bench.py --variant U_to_gradP --steps 4 --warmup 3 --no-cpu-baseline --no-c5"
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:tc_gemm_grid|pca_inverse_t|task_means|reduce_standardise|offsets_kernel' -s 20 -c 5 -f -o $OUT/prof_${TAG}_c3 $CMD3 > $OUT/ncu_full_${TAG}_c3.log 2>&1
echo "c3 capture rc=$?"
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_${TAG}_*.json')):
    try:
        b=json.loads(open(f).read().strip().splitlines()[-1])
        e=b['e2e']; c5=b.get('roofline_c5') or {}
        print(f.split('/')[-1], 'ms', round(b['ms_per_step'],4), 'cells/s %.3g' % b['value'], 'e2e', round(e['ms_per_step'],4), {k:round(v['ms']*1e3,1) for k,v in b['stages'].items() if v['ms']>0.003}, 'c5', c5.get('ms_per_step'), b.get('latency'))
    except Exception as ex: print(f, 'ERR', ex)
PY
