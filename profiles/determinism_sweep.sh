#!/bin/bash
# which switch removes the step-to-step nondeterminism seen under programmatic dependent launch?
for v in NONE PSM_TF32_MASK_HI PSM_NO_FUSED_EXTRACT PSM_NO_PREP_BULK PSM_NO_INV_T PSM_NO_FUSED_OFFSETS PSM_NO_FUSED_PLACE PSM_NO_DENSE_CLUSTER PSM_NO_GRAPHS; do
  echo "== $v"; env $v=1 timeout 150 python profiles/determinism_probe.py c2 2>&1 | grep "call [15]" | cut -c1-200
done
