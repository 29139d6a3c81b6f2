set -u
timeout 900 python -m pytest tests/test_gpu_ik.py tests/test_gpu_round2.py -m gpu -x -q --timeout=900 2>&1 | tail -3
python tools/dev/dev_ik_time.py 20 22 24 2>&1 | tail -6
