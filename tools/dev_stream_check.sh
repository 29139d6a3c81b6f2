#!/bin/bash
for s in 1 0; do
  echo "== PNP_IK_STREAM=$s"
  PNP_IK_STREAM=$s python tools/dev_pair_check.py 12 18 20 22 24 2>&1 | grep -E "ik f32 spec|bitwise: False|per-query"
done
python -m pytest tests/test_gpu_ik.py -x -q 2>&1 | tail -3
