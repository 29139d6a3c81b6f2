"""Developer scratch: the planner at 2^20 reachable plans, a few launches (ncu target)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from mujoco_panda_pnp_b200 import engine, synthetic, KinematicTree
tree = KinematicTree.from_mjcf(); engine.set_tree(tree)
dev = torch.device("cuda")
n = 1 << int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
wp = synthetic.reachable_move_envs(n, tree.lower, tree.upper, seed=1, device=dev)
goal = engine.fk_jac(wp["q_goal"], want_quat=False, want_jac=False)[0]
pk = engine.ik_params()
cnt = torch.zeros(4, dtype=torch.int64, device=dev)
out = engine.move_ik_plan(wp["q_start"], goal, pk, counters=cnt, traj_cap=256)
torch.cuda.synchronize()
c = cnt.cpu().numpy()
print(f"solves {c[0]} conv {c[1]} mean evaluations per solve {c[3] / c[0]:.3f}, solves per plan {c[0] / n:.2f}")
for _ in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); engine.move_ik_plan(wp["q_start"], goal, pk, traj_cap=256, out=out); e1.record(); torch.cuda.synchronize()
    print(f"{e0.elapsed_time(e1):.3f} ms")
