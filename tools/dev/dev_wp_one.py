"""Developer scratch: cfg4 (2^20 envs x 50 waypoints), a few launches (ncu target)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from mujoco_panda_pnp_b200 import engine, synthetic, KinematicTree
tree = KinematicTree.from_mjcf(); engine.set_tree(tree)
dev = torch.device("cuda")
n = 1 << 20
w = synthetic.waypoint_envs(n, seed=0, device=dev)
p = engine.ik_params(kinematics=sys.argv[1] if len(sys.argv) > 1 else "auto")
for _ in range(4):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); engine.ik_waypoints(w["q_start"], w["goal"], 50, p); e1.record(); torch.cuda.synchronize()
    print(f"{e0.elapsed_time(e1):.3f} ms")
