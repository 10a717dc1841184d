#!/bin/bash
# Round-2 evidence capture (run under gpurun, one GPU):  bash profiles/capture_r2.sh <tag> [notest] [noncu]
# 1. GPU parity tests, 2. bench lines (c2 default, c5, c3), 3. ncu launch list of the bench command,
# 4. one `ncu --set full` capture of every kernel of one step.  Outputs land in gpurun_out/.
TAG=${1:-r2}
OUT=gpurun_out
mkdir -p $OUT
if [ "$2" != "notest" ]; then
python -m pytest tests -m gpu -x -q > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_$TAG.log
tail -5 $OUT/pytest_$TAG.log
fi
python bench.py > $OUT/bench_${TAG}_c2.json 2> $OUT/bench_${TAG}_c2.err; echo "bench c2 rc=$?"
python bench.py --workload c5 --no-cpu-baseline > $OUT/bench_${TAG}_c5.json 2> $OUT/bench_${TAG}_c5.err; echo "bench c5 rc=$?"
python bench.py --variant U_to_gradP --no-cpu-baseline > $OUT/bench_${TAG}_c3.json 2> $OUT/bench_${TAG}_c3.err; echo "bench c3 rc=$?"
if [ "$3" != "noncu" ]; then
K='regex:^(void )?(psm::)?(prep|gather|extract|tc_gemm|dense_cluster|dense_stack|pca_inverse_t|reduce_standardise|task_|offsets|place|back|gauss|p2p|init_)'
CMD="python bench.py --steps 4 --warmup 3 --no-cpu-baseline"
$CMD > $OUT/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches_$TAG.csv $CMD > $OUT/ncu_launch_$TAG.log 2>&1
echo "launch list rc=$?"
$CMD > $OUT/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k "$K" -s ${NCU_SKIP:-44} -c ${NCU_COUNT:-11} -f -o $OUT/prof_$TAG $CMD > $OUT/ncu_full_$TAG.log 2>&1
echo "full capture rc=$?"
fi
cat $OUT/bench_${TAG}_c2.json
