"""One cfg2 launch (4096 cold targets) after a warm-up, for `ncu -k regex:ik_solve_small_kernel` (development tool)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from mujoco_panda_pnp_b200 import KinematicTree, engine, synthetic  # noqa: E402

tree = KinematicTree.from_mjcf()
engine.set_tree(tree)
kin = sys.argv[1] if len(sys.argv) > 1 else "auto"
neutral = torch.tensor([0.00, 0.41, 0.00, -1.85, 0.00, 2.26, 0.79], dtype=torch.float32, device="cuda")
q = synthetic.random_joint_configs(4096, tree.lower, tree.upper, seed=1234, device="cuda")
t = engine.fk_jac(q, want_quat=False, want_jac=False)[0]
for _ in range(3):
    r = engine.ik_solve(t, neutral, engine.ik_params(kinematics=kin))
torch.cuda.synchronize()
print(int(r.converged.sum()), int(r.iterations.max()))
