#!/bin/bash
# round 2, call d (1 GPU): A/B of kernel variants on c2 and c5 (same box), one bench line each
OUT=gpurun_out; mkdir -p $OUT
B="python bench.py --no-c5 --no-cpu-baseline --steps 50"
$B > $OUT/ab_default_c2.json 2> $OUT/ab.err; echo "default rc=$?"
PSM_GATHER_U=1 $B > $OUT/ab_gu1_c2.json 2>> $OUT/ab.err; echo "gu1 rc=$?"
PSM_BACK_SCALAR=1 $B > $OUT/ab_backscalar_c2.json 2>> $OUT/ab.err; echo "backscalar rc=$?"
PSM_NO_STRIP_FUSE=1 $B > $OUT/ab_nostrip_c2.json 2>> $OUT/ab.err; echo "nostrip rc=$?"
PSM_PREP_BULK=0 $B > $OUT/ab_noprepbulk_c2.json 2>> $OUT/ab.err; echo "noprepbulk rc=$?"
PSM_NO_PREFETCH=1 $B > $OUT/ab_noprefetch_c2.json 2>> $OUT/ab.err; echo "noprefetch rc=$?"
B5="python bench.py --workload c5 --no-cpu-baseline --steps 30"
$B5 > $OUT/ab_default_c5.json 2>> $OUT/ab.err; echo "default c5 rc=$?"
PSM_GATHER_U=1 $B5 > $OUT/ab_gu1_c5.json 2>> $OUT/ab.err; echo "gu1 c5 rc=$?"
PSM_NO_STRIP_FUSE=1 $B5 > $OUT/ab_nostrip_c5.json 2>> $OUT/ab.err; echo "nostrip c5 rc=$?"
python bench.py --no-c5 --steps 20 --variant U_to_gradP --no-cpu-baseline > $OUT/ab_default_c3.json 2>> $OUT/ab.err; echo "c3 rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/ab_*.json')):
    try:
        b=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], 'ms', round(b['ms_per_step'],4), 'e2e', round(b['e2e']['ms_per_step'],4), {k:round(v['ms']*1e3,1) for k,v in b['stages'].items() if v['ms']>0.003})
    except Exception as e: print(f, 'ERR', e)
PY
tail -5 $OUT/ab.err
