#!/bin/bash
# sweep the deferred-flush threshold of ik_solve_v_kernel (developer tool)
for f in ${FLUSHES:-4 8 12}; do
  echo "== PNP_IK_FLUSH_MIN=$f"
  PNP_IK_FLUSH_MIN=$f python tools/dev/dev_pair_check.py 24 2>&1 | grep -E "ik f32|bitwise: False"
done
echo "== default"
python tools/dev/dev_pair_check.py 12 18 20 22 2>&1 | grep -E "ik f32|bitwise: False|per-query"
