"""Developer scratch: what bounds the planner?  A = as shipped (longest plan first through `order`), B = inputs gathered
into plan order on the host and taken in index order (no order[] indirection, sequential rows), C = trajectory
capacity 1 (the appends are predicated off), B+C both."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from mujoco_panda_pnp_b200 import engine, synthetic, KinematicTree
tree = KinematicTree.from_mjcf(); engine.set_tree(tree)
dev = torch.device("cuda")
n = 1 << 20
wp = synthetic.reachable_move_envs(n, tree.lower, tree.upper, seed=1, device=dev)
goal = engine.fk_jac(wp["q_goal"], want_quat=False, want_jac=False)[0]
pk = engine.ik_params()
order = engine.move_plan_order(wp["q_start"], goal)
qs_sorted = wp["q_start"][order.long()].contiguous(); goal_sorted = goal[order.long()].contiguous()

def run(name, qs, gl, cap, order):
    out = engine.move_ik_plan(qs, gl, pk, traj_cap=cap, order=order)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); engine.move_ik_plan(qs, gl, pk, traj_cap=cap, out=out, order=order); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print(f"{name:60s} {min(ts):.3f} ms  {n / min(ts) / 1e3:.0f} M plans/s", flush=True)

run("A  auto (sorted records in the call), cap 256", wp["q_start"], goal, 256, "auto")
run("A' order tensor precomputed (indirect reads), cap 256", wp["q_start"], goal, 256, order)
run("B  inputs pre-gathered on the host, index order, cap 256", qs_sorted, goal_sorted, 256, None)
run("C  auto, cap 2 (hardly any appends; direct-store kernel)", wp["q_start"], goal, 2, "auto")
run("D  index order (no longest-first), cap 256", wp["q_start"], goal, 256, None)
run("E  auto, cap 128", wp["q_start"], goal, 128, "auto")
