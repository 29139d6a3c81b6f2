#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q --timeout=900 > $O/r2g_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/r2g_tests.log
timeout 600 python tools/dev/dev_ik_fixed_cost.py > $O/r2g_fixed.json 2>/dev/null; echo "rc=$?"
for F in 8 12; do PNP_IK_FLUSH_MIN=$F timeout 600 python tools/dev/dev_ik_fixed_cost.py > $O/r2g_fixed_flush$F.json 2>/dev/null; done
python - <<'PY'
import json
for t in ("","_flush8","_flush12"):
    d=json.load(open(f"gpurun_out/r2g_fixed{t}.json"))
    print(t, {k:(x.get("fixed_ms"),x.get("ns_per_query")) for k,x in d.items() if "fixed_ms" in x}, d["cold/100"]["ms"]["16777216"])
PY
