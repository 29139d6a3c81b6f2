"""Developer scratch: does longest-plan-first ordering shorten the planner's tail?  (inputs physically permuted with torch)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from mujoco_panda_pnp_b200 import engine, synthetic, KinematicTree
tree = KinematicTree.from_mjcf(); engine.set_tree(tree)
dev = torch.device("cuda")
def timeit(fn, warm=1, rep=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(rep):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts)
for logn in (14, 16, 18, 20):
    n = 1 << logn
    wp = synthetic.reachable_move_envs(n, tree.lower, tree.upper, seed=1, device=dev)
    goal = engine.fk_jac(wp["q_goal"], want_quat=False, want_jac=False)[0]
    p0 = engine.fk_jac(wp["q_start"], want_quat=False, want_jac=False)[0]
    d0 = (goal - p0).norm(dim=1)
    pk = engine.ik_params(kinematics="spec_lane")
    out = engine.move_ik_plan(wp["q_start"], goal, pk, traj_cap=128)
    ns = out["n_solves"].float()
    print(f"2^{logn}: corr(d0, n_solves) = {torch.corrcoef(torch.stack([d0, ns]))[0,1].item():.3f}; solves mean {ns.mean().item():.1f} min {ns.min().item():.0f} max {ns.max().item():.0f}")
    for name, perm in (("index order", None), ("d0 descending", torch.argsort(d0, descending=True)),
                       ("d0 desc, 64 buckets", torch.argsort((d0 * 32).int().clamp(max=63), descending=True)),
                       ("n_solves descending (ideal)", torch.argsort(ns, descending=True))):
        qs = wp["q_start"] if perm is None else wp["q_start"][perm].contiguous()
        g = goal if perm is None else goal[perm].contiguous()
        best = timeit(lambda: engine.move_ik_plan(qs, g, pk, traj_cap=128))
        print(f"  {name:30s}: {best:.3f} ms -> {n / best / 1e3:.1f} M plans/s")
    del wp, goal, out
