set -u
for rep in 1 2; do
for V in "-DIK_CITER64=1" ""; do
  PNP_NVCC_EXTRA="$V" python -m mujoco_panda_pnp_b200.csrc.build --force > /dev/null 2>&1 || { echo build failed; exit 1; }
  echo "== variant '$V'"; python tools/dev/dev_ik_time.py 24 2>&1 | grep "ik 2"
done
done
