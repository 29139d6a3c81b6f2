"""Developer scratch: pose-mode IK timing (bench side entry workload)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from mujoco_panda_pnp_b200 import engine, synthetic, KinematicTree
tree = KinematicTree.from_mjcf(); engine.set_tree(tree)
dev = torch.device("cuda")
n_p = 1 << 20
qp = synthetic.reachable_move_envs(n_p, tree.lower, tree.upper, seed=5, device=dev, spread=0.5)["q_goal"]
ppos, pquat, _ = engine.fk_jac(qp, want_jac=False)
neutral = torch.tensor(synthetic.NEUTRAL_Q, device=dev)
cnt = torch.zeros(4, dtype=torch.int64, device=dev)
p = engine.ik_params()
f = lambda c=None: engine.ik_pose_solve(ppos, pquat, neutral, p, counters=c)
f(cnt); torch.cuda.synchronize(); c = cnt.cpu().numpy()
for _ in range(2): f()
torch.cuda.synchronize(); ts = []
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); f(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
ms = min(ts)
print(f"pose ik 2^20: {ms:.4f} ms -> {n_p / ms / 1e6:.3f} G solves/s, conv {c[1] / c[0]:.4f}, mean it {c[3] / c[0]:.3f}, {950.0 * c[3] / ms / 1e9:.1f} TFLOP/s algorithmic")
