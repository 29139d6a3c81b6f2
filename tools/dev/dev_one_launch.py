"""One IK launch per kinematics selector given on the command line (for ncu captures)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from mujoco_panda_pnp_b200 import engine, synthetic, KinematicTree
tree = KinematicTree.from_mjcf()
engine.set_tree(tree)
dev = torch.device("cuda")
logn = int(sys.argv[1]); kins = sys.argv[2:]
n = 1 << logn
neutral = torch.tensor(synthetic.NEUTRAL_Q, device=dev)
qs = synthetic.random_joint_configs(n, tree.lower, tree.upper, seed=1234, device=dev)
tg, _, _ = engine.fk_jac(qs, want_quat=False, want_jac=False)
for kin in kins:
    for _ in range(3):
        engine.ik_solve(tg, neutral, engine.ik_params(kinematics=kin))
torch.cuda.synchronize()
print("ok")
