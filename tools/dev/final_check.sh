#!/bin/bash
# everything once on the GPU box: GPU tests, smoke(), the default bench line
set -u
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q --timeout=900 2>&1 | tail -4
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python bench.py --steps 20 --warmup 5 > $O/r2_final_bench.json 2> $O/r2_final_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_final_bench.json').read().strip().splitlines()[-1])
print(json.dumps(d['headline']))
print(d['ms_per_step'], d['roofline']['frac'])
PY
