"""Developer scratch: cfg4 waypoint kernel variants."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from mujoco_panda_pnp_b200 import engine, synthetic, KinematicTree
tree = KinematicTree.from_mjcf(); engine.set_tree(tree)
dev = torch.device("cuda")
def timeit(fn, warm=2, rep=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(rep):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts)
n = 1 << 20
w = synthetic.waypoint_envs(n, seed=0, device=dev)
for kin in ("generic", "spec_lane", "spec_pair"):
    p = engine.ik_params(kinematics=kin)
    cnt = torch.zeros(4, dtype=torch.int64, device=dev)
    r = engine.ik_waypoints(w["q_start"], w["goal"], 50, p, counters=cnt); torch.cuda.synchronize()
    c = cnt.cpu().numpy()
    best = timeit(lambda: engine.ik_waypoints(w["q_start"], w["goal"], 50, p))
    print(f"waypoints {kin:10s} n=2^20 x50: best {best:.3f} ms -> {c[0] / best / 1e6:.2f} G warm solves/s; mean it {c[3] / c[0]:.3f} conv {c[1] / c[0]:.4f} accepted mean {r['n_accepted'].float().mean().item():.2f}")
