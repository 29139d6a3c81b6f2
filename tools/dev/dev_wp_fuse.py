"""Developer scratch: fused accept + first-iteration pass of the waypoint kernels: bit identity and time."""
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
for n in (1000, 1 << 20):
    w = synthetic.waypoint_envs(n, seed=0, device=dev)
    for kin in ("spec_lane", "spec_pair"):
        p = engine.ik_params(kinematics=kin)
        res = {}
        for fuse in ("0", "1"):
            os.environ["PNP_WAYPOINT_FUSE"] = fuse
            cnt = torch.zeros(4, dtype=torch.int64, device=dev)
            r = engine.ik_waypoints(w["q_start"], w["goal"], 50, p, counters=cnt); torch.cuda.synchronize()
            best = timeit(lambda: engine.ik_waypoints(w["q_start"], w["goal"], 50, p))
            c = cnt.cpu().numpy()
            res[fuse] = (r, cnt.clone())
            print(f"n={n} {kin} fuse={fuse}: {best:.3f} ms -> {c[0] / best / 1e6:.2f} G warm solves/s; mean it {c[3] / c[0]:.3f}")
        (a, ca), (b, cb) = res["0"], res["1"]
        same = all(torch.equal(a[f], b[f]) for f in a.keys() if isinstance(a[f], torch.Tensor)) and torch.equal(ca, cb)
        print(f"   bit-identical: {same}", {f: float((a[f].double() - b[f].double()).abs().max()) for f in a.keys() if isinstance(a[f], torch.Tensor)})
