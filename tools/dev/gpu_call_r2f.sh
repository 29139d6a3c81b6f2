#!/bin/bash
# round 2, GPU call F: tail phase x guided ticket chunks
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q --timeout=900 > $O/r2f_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/r2f_tests.log
for T in 1 0; do for G in 1 0; do
PNP_IK_TAIL=$T PNP_IK_GUIDED=$G timeout 600 python tools/dev/dev_ik_fixed_cost.py > $O/r2f_fixed_t${T}g${G}.json 2>/dev/null; echo "t$T g$G rc=$?"
done; done
python - <<'PY'
import json
for t in ("t1g1","t1g0","t0g1","t0g0"):
    d=json.load(open(f"gpurun_out/r2f_fixed_{t}.json"))
    print(t, {k:(x.get("fixed_ms"),x.get("ns_per_query")) for k,x in d.items() if "fixed_ms" in x}, d["cold/100"]["ms"]["16777216"])
PY
