set -u
timeout 900 python -m pytest tests/test_gpu_ik.py tests/test_gpu_round2.py -m gpu -x -q --timeout=900 2>&1 | tail -3
python tools/dev/dev_ik_time.py 20 22 24 26 2>&1 | tail -8
for f in 6 8 12; do echo "== FLUSH_MIN $f"; PNP_IK_FLUSH_MIN=$f python tools/dev/dev_ik_time.py 24 2>&1 | tail -2 | head -1; done
