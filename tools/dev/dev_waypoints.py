"""Developer scratch: time the cfg4 waypoint kernel and the planner."""
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
p = engine.ik_params()
cnt = torch.zeros(4, dtype=torch.int64, device=dev)
r = engine.ik_waypoints(w["q_start"], w["goal"], 50, p, counters=cnt); torch.cuda.synchronize()
c = cnt.cpu().numpy()
best = timeit(lambda: engine.ik_waypoints(w["q_start"], w["goal"], 50, p))
print(f"waypoints n=2^20 x50: best {best:.3f} ms -> {c[0] / best / 1e6:.2f} G warm solves/s; mean it {c[3] / c[0]:.3f} conv {c[1] / c[0]:.4f} accepted mean {r['n_accepted'].float().mean().item():.2f}")

# planner: reachable goals, and a batch with 1/16 unreachable goals mixed in
import numpy as np
n_pl = 1 << 18
wp = synthetic.reachable_move_envs(n_pl, tree.lower, tree.upper, seed=1, device=dev)
goal = engine.fk_jac(wp["q_goal"], want_quat=False, want_jac=False)[0]
cnt = torch.zeros(4, dtype=torch.int64, device=dev)
engine.move_ik_plan(wp["q_start"], goal, p, counters=cnt); torch.cuda.synchronize()
cp = cnt.cpu().numpy()
best = timeit(lambda: engine.move_ik_plan(wp["q_start"], goal, p), warm=1, rep=3)
print(f"planner reachable 2^18: best {best:.3f} ms -> {n_pl / best / 1e3:.1f} M plans/s, {cp[0] / best / 1e6:.2f} G solves/s, mean it {cp[3] / cp[0]:.2f}")
n_mx = 1 << 14
goal_mx = goal[:n_mx].clone(); goal_mx[::16] = torch.tensor([2.5, 0.0, 0.5], device=dev)
cnt.zero_(); out = engine.move_ik_plan(wp["q_start"][:n_mx].contiguous(), goal_mx, p, counters=cnt); torch.cuda.synchronize()
cp = cnt.cpu().numpy()
best = timeit(lambda: engine.move_ik_plan(wp["q_start"][:n_mx].contiguous(), goal_mx, p), warm=1, rep=3)
print(f"planner mixed (1/16 unreachable) 2^14: best {best:.3f} ms -> {n_mx / best / 1e3:.2f} M plans/s, {cp[0] / best / 1e6:.2f} G solves/s, mean it {cp[3] / cp[0]:.2f}, status!=0: {(out['status'] != 0).sum().item()}")
n_mx = 1 << 18
goal_mx = goal.clone(); goal_mx[::64] = torch.tensor([2.5, 0.0, 0.5], device=dev)
qsm = wp["q_start"]
cnt.zero_(); out = engine.move_ik_plan(qsm, goal_mx, p, counters=cnt); torch.cuda.synchronize()
cp = cnt.cpu().numpy()
best = timeit(lambda: engine.move_ik_plan(qsm, goal_mx, p), warm=0, rep=2)
print(f"planner mixed (1/64 unreachable) 2^18: best {best:.3f} ms -> {n_mx / best / 1e3:.2f} M plans/s, {cp[0] / best / 1e6:.2f} G solves/s, mean it {cp[3] / cp[0]:.2f}, status!=0: {(out['status'] != 0).sum().item()}")
for kin in ("generic", "spec_lane", "spec_pair"):
    pk = engine.ik_params(kinematics=kin)
    cnt.zero_(); engine.move_ik_plan(wp["q_start"], goal, pk, counters=cnt); torch.cuda.synchronize()
    cp = cnt.cpu().numpy()
    best = timeit(lambda: engine.move_ik_plan(wp["q_start"], goal, pk), warm=1, rep=3)
    print(f"planner {kin:10s} reachable 2^18: best {best:.3f} ms -> {n_pl / best / 1e3:.1f} M plans/s, {cp[0] / best / 1e6:.2f} G solves/s, mean it {cp[3] / cp[0]:.2f}")
