#!/bin/bash
# round 2, GPU call C: fixed-cost decomposition of the pair kernel + regression check after dropping the hand-over
set -u
O=gpurun_out
mkdir -p $O
timeout 600 python tools/dev/dev_ik_fixed_cost.py > $O/r2c_fixed_cost.json 2> $O/r2c_fixed_cost.err; echo "fixed rc=$?"; cat $O/r2c_fixed_cost.json
timeout 600 python tools/dev/dev_ik_variants.py --quick > $O/r2c_variants.json 2> $O/r2c_variants.err; echo "variants rc=$?"
for OCC in 2 3; do PNP_IK_OCC=$OCC timeout 300 python tools/dev/dev_ik_fixed_cost.py > $O/r2c_fixed_cost_occ$OCC.json 2>/dev/null; echo "occ$OCC rc=$?"; done
timeout 600 python -m pytest tests -m gpu -q --timeout=900 > $O/r2c_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/r2c_tests.log
