#!/bin/bash
# round 2, call g (one GPU): GPU parity tests, the default bench line (c2 + roofline_c5 + cpu_baseline), c3, A/B of the chunked
# host tail (PSM_TAIL_CHUNKS), then the ncu launch list and one full capture of a step.
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r2g}
python -m pytest tests -m gpu -x -q > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_$TAG.log
python bench.py > $OUT/bench_${TAG}_c2.json 2> $OUT/bench_${TAG}_c2.err; echo "bench c2 rc=$?"
python bench.py --variant U_to_gradP --no-cpu-baseline > $OUT/bench_${TAG}_c3.json 2> $OUT/bench_${TAG}_c3.err; echo "bench c3 rc=$?"
for c in 1 2 8; do
PSM_TAIL_CHUNKS=$c python bench.py --no-cpu-baseline --no-c5 --steps 30 > $OUT/ab_${TAG}_tail$c.json 2>> $OUT/ab_$TAG.err; echo "tail $c rc=$?"
done
python profiles/pcie_probe.py > $OUT/pcie_$TAG.log 2>&1
K='regex:^(void )?(psm::)?(prep|gather|extract|tc_gemm|dense_cluster|dense_stack|pca_inverse_t|reduce_standardise|task_|offsets|place|back|gauss|p2p|init_)'
CMD="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-c5"
$CMD > $OUT/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches_$TAG.csv $CMD > $OUT/ncu_launch_$TAG.log 2>&1
echo "launch list rc=$?"
$CMD > $OUT/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k "$K" -s ${NCU_SKIP:-44} -c ${NCU_COUNT:-11} -f -o $OUT/prof_$TAG $CMD > $OUT/ncu_full_$TAG.log 2>&1
echo "full capture rc=$?"
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/*_${TAG}_*.json')):
    try:
        b=json.loads(open(f).read().strip().splitlines()[-1])
        e=b['e2e']
        print(f.split('/')[-1], 'ms', round(b['ms_per_step'],4), 'e2e', round(e['ms_per_step'],4), 'rows5', round(e.get('rows5',{}).get('ms_per_step',0),4), 'cold', round(e.get('cold_host_cache',{}).get('ms_per_step',0),4), {k:round(v['ms']*1e3,1) for k,v in b['stages'].items() if v['ms']>0.003})
    except Exception as ex: print(f, 'ERR', ex)
PY
cat $OUT/pcie_$TAG.log
