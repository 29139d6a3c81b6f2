set -u
for T in 8 16 32; do
  PNP_NVCC_EXTRA="-DIK_HANDOVER_AT=$T" python -m mujoco_panda_pnp_b200.csrc.build --force > /dev/null 2>&1 || { echo build failed; exit 1; }
  echo "== handover at <= $T running slots per warp"; python tools/dev/dev_ik_time.py 20 22 24 2>&1 | grep "ik 2"
done
python -m mujoco_panda_pnp_b200.csrc.build --force > /dev/null 2>&1
timeout 900 python -m pytest tests/test_gpu_ik.py tests/test_gpu_round2.py -m gpu -x -q --timeout=900 2>&1 | tail -2
