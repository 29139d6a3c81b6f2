#!/bin/bash
# round 2, GPU call A: tests, variant timings, smoke, bench, one ncu capture of the latency kernel
set -u
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > $O/r2a_gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 > $O/r2a_tests.log 2>&1; echo "tests rc=$?" | tee -a $O/r2a_tests.log
tail -5 $O/r2a_tests.log
timeout 600 python tools/dev/dev_ik_variants.py > $O/r2a_variants_park1.json 2> $O/r2a_variants_park1.err; echo "variants rc=$?"
PNP_IK_PARK=0 timeout 300 python tools/dev/dev_ik_variants.py --quick > $O/r2a_variants_park0.json 2> $O/r2a_variants_park0.err; echo "variants park0 rc=$?"
PNP_IK_FLUSH_MIN=8 timeout 300 python tools/dev/dev_ik_variants.py --quick > $O/r2a_variants_flush8.json 2>/dev/null; echo "flush8 rc=$?"
PNP_IK_FLUSH_MIN=13 timeout 300 python tools/dev/dev_ik_variants.py --quick > $O/r2a_variants_flush13.json 2>/dev/null; echo "flush13 rc=$?"
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2a_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $O/r2a_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > $O/r2a_bench.json 2> $O/r2a_bench.err; echo "bench rc=$?"; tail -c 1500 $O/r2a_bench.json
if timeout 120 python tools/dev/dev_cfg2_once.py auto > $O/r2a_cfg2_plain.log 2>&1; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:ik_solve_small_kernel -s 2 -c 1 -f -o $O/r2a_small python tools/dev/dev_cfg2_once.py auto > $O/r2a_ncu_small.log 2>&1; echo "ncu small rc=$?"
fi
ls -la $O | tail -20
