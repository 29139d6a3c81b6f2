"""Developer scratch: fk_jac kernel timing."""
import os, sys, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from mujoco_panda_pnp_b200 import engine, synthetic, KinematicTree
tree = KinematicTree.from_mjcf(); engine.set_tree(tree)
dev = torch.device("cuda")
n = 1 << 22
q = synthetic.random_joint_configs(n, tree.lower, tree.upper, seed=9, device=dev)
for kw in (dict(), dict(want_jac=False), dict(want_quat=False, want_jac=False)):
    f = lambda: engine.fk_jac(q, **kw)
    for _ in range(3): f()
    torch.cuda.synchronize(); ts = []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ms = statistics.median(ts)
    b = 28 + 12 + (16 if kw.get("want_quat", True) else 0) + (168 if kw.get("want_jac", True) else 0)
    print(f"fk_jac {kw}: {ms:.4f} ms -> {n / ms / 1e6:.2f} G configs/s, {b * n / ms / 1e6:.0f} GB/s ({b} B/config)")
