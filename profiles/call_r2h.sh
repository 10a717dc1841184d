#!/bin/bash
# round 2, call h (one GPU): one-launch projection (cluster fold) + pre-split Dense operands + two-stream uploads: parity tests
# first (bounded: a wrong barrier must not hang the box), then bench lines and the A/B switches, then the PCIe probe.
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r2h}
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q > $OUT/pytest_${TAG}_quick.log 2>&1; rc=$?; echo "quick rc=$rc"; tail -3 $OUT/pytest_${TAG}_quick.log
if [ $rc -ne 0 ]; then tail -40 $OUT/pytest_${TAG}_quick.log; exit 1; fi
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_$TAG.log
timeout 600 python bench.py --no-cpu-baseline > $OUT/bench_${TAG}_c2.json 2> $OUT/bench_${TAG}_c2.err; echo "bench c2 rc=$?"
timeout 600 python bench.py --variant U_to_gradP --no-cpu-baseline > $OUT/bench_${TAG}_c3.json 2> $OUT/bench_${TAG}_c3.err; echo "bench c3 rc=$?"
for sw in PSM_NO_PROJ_CLUSTER PSM_NO_DENSE_PRESPLIT PSM_NO_SPLIT_H2D; do
env $sw=1 timeout 300 python bench.py --no-cpu-baseline --no-c5 --steps 30 > $OUT/ab_${TAG}_$sw.json 2>> $OUT/ab_$TAG.err; echo "$sw rc=$?"
done
PSM_PROJ_KS=4 timeout 300 python bench.py --no-cpu-baseline --no-c5 --steps 30 > $OUT/ab_${TAG}_ks4.json 2>> $OUT/ab_$TAG.err; echo "ks4 rc=$?"
timeout 300 python profiles/pcie_probe.py > $OUT/pcie_$TAG.log 2>&1
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/*_${TAG}_*.json')):
    try:
        b=json.loads(open(f).read().strip().splitlines()[-1])
        e=b['e2e']
        print(f.split('/')[-1], 'ms', round(b['ms_per_step'],4), 'e2e', round(e['ms_per_step'],4), 'rows5', round(e.get('rows5',{}).get('ms_per_step',0),4), {k:round(v['ms']*1e3,1) for k,v in b['stages'].items() if v['ms']>0.003})
    except Exception as ex: print(f, 'ERR', ex)
PY
cat $OUT/pcie_$TAG.log; tail -5 $OUT/ab_$TAG.err
