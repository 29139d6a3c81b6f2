#!/bin/bash
# quick A/B of the pair kernel on the GPU box: fixed-cost fit at the default and two other flush thresholds
set -u
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_ik.py tests/test_gpu_round2.py tests/test_gpu_move.py -m gpu -q --timeout=900 2>&1 | tail -2
for F in 0 6 8 12; do
  if [ $F = 0 ]; then timeout 300 python tools/dev/dev_ik_fixed_cost.py > $O/ab_fixed.json 2>/dev/null; T=""; else PNP_IK_FLUSH_MIN=$F timeout 300 python tools/dev/dev_ik_fixed_cost.py > $O/ab_fixed_f$F.json 2>/dev/null; T="_f$F"; fi
  python -c "
import json; d=json.load(open('$O/ab_fixed$T.json')); print('flush_min=$F', {k:(x.get('fixed_ms'),x.get('ns_per_query')) for k,x in d.items() if 'fixed_ms' in x}, d['cold/100']['ms']['16777216'])"
done
