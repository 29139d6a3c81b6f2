set -u
timeout 900 python -m pytest tests -m gpu -x -q --timeout=900 2>&1 | tail -8
python tools/dev/dev_ik_time.py 20 22 24 2>&1 | tail -6
python tools/dev/dev_planner_exp.py 2>&1 | head -2
