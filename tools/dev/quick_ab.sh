#!/bin/bash
# quick check of the pair kernel on the GPU box: IK / round-2 / planner tests, then the fixed-cost fit
set -u
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_ik.py tests/test_gpu_round2.py tests/test_gpu_move.py tests/test_gpu_host_api.py -m gpu -q --timeout=900 2>&1 | tail -2
timeout 300 python tools/dev/dev_ik_fixed_cost.py > $O/ab_fixed.json 2>/dev/null
python -c "
import json; d=json.load(open('$O/ab_fixed.json')); print({k:(x.get('fixed_ms'),x.get('ns_per_query')) for k,x in d.items() if 'fixed_ms' in x}, d['cold/100']['ms'])"
