"""Developer scratch: planner throughput vs batch size (load imbalance at small batches)."""
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
    for kin, order, fuse in (("spec_lane", "auto", "0"), ("spec_lane", "auto", "1"), ("spec_pair", "auto", "1")):
        pk = engine.ik_params(kinematics=kin); os.environ["PNP_WAYPOINT_FUSE"] = fuse
        cnt = torch.zeros(4, dtype=torch.int64, device=dev)
        out = engine.move_ik_plan(wp["q_start"], goal, pk, counters=cnt, traj_cap=128); torch.cuda.synchronize()
        cp = cnt.cpu().numpy()
        best = timeit(lambda: engine.move_ik_plan(wp["q_start"], goal, pk, traj_cap=128, order=order, out=out))
        print(f"planner {kin} order={order} fuse={fuse} 2^{logn}: {best:.3f} ms -> {n / best / 1e3:.1f} M plans/s, {cp[0] / best / 1e6:.2f} G solves/s, max len {int(out['traj_len'].max())}")
    del wp, goal, out
