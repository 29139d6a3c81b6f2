#!/bin/bash
# round 2, GPU call B: tests, pair-kernel A/B (hand-over on/off), smoke, bench
set -u
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 > $O/r2b_tests.log 2>&1; echo "tests rc=$?" | tee -a $O/r2b_tests.log
tail -5 $O/r2b_tests.log
timeout 600 python tools/dev/dev_ik_variants.py > $O/r2b_variants_park1.json 2> $O/r2b_variants_park1.err; echo "variants rc=$?"
PNP_IK_PARK=0 timeout 300 python tools/dev/dev_ik_variants.py --quick > $O/r2b_variants_park0.json 2> $O/r2b_variants_park0.err; echo "variants park0 rc=$?"
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2b_smoke.log 2>&1; echo "smoke rc=$?"; tail -4 $O/r2b_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > $O/r2b_bench.json 2> $O/r2b_bench.err; echo "bench rc=$?"; tail -c 1200 $O/r2b_bench.json
