#!/bin/bash
# round 2, GPU call E: pair kernel with the tail phase on / off
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q --timeout=900 > $O/r2e_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/r2e_tests.log
timeout 600 python tools/dev/dev_ik_fixed_cost.py > $O/r2e_fixed_cost_tail1.json 2> $O/r2e_fixed_cost.err; echo "fixed rc=$?"
PNP_IK_TAIL=0 timeout 600 python tools/dev/dev_ik_fixed_cost.py > $O/r2e_fixed_cost_tail0.json 2>/dev/null; echo "fixed0 rc=$?"
timeout 600 python tools/dev/dev_ik_variants.py --quick > $O/r2e_variants_tail1.json 2>/dev/null; echo "variants rc=$?"
PNP_IK_TAIL=0 timeout 600 python tools/dev/dev_ik_variants.py --quick > $O/r2e_variants_tail0.json 2>/dev/null; echo "variants0 rc=$?"
python - <<'PY'
import json
for t in ("tail1","tail0"):
    d=json.load(open(f"gpurun_out/r2e_fixed_cost_{t}.json")); v=json.load(open(f"gpurun_out/r2e_variants_{t}.json"))
    print(t, {k:(x.get("fixed_ms"),x.get("ns_per_query")) for k,x in d.items() if "fixed_ms" in x}, v["ik_2^24_spec_pair"])
PY
