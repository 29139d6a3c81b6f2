"""Developer scratch: the bench's IK workload (cold reachable targets, broadcast q_init, packed output) at 2^LOG queries,
timed as K back-to-back launches between one pair of events (what bench.py's timed region does) and one by one."""
import os, sys, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from mujoco_panda_pnp_b200 import KinematicTree, engine, synthetic
tree = KinematicTree.from_mjcf(); engine.set_tree(tree)
dev = torch.device("cuda")
neutral = torch.tensor(synthetic.NEUTRAL_Q, dtype=torch.float32, device=dev)
for lg in [int(x) for x in (sys.argv[1:] or ["24"])]:
    n = 1 << lg
    tg = torch.empty((n, 3), device=dev)
    for off in range(0, n, 1 << 22):
        m = min(1 << 22, n - off)
        q = synthetic.random_joint_configs(m, tree.lower, tree.upper, seed=1234 + off, device=dev)
        tg[off:off + m] = engine.fk_jac(q, want_quat=False, want_jac=False)[0]
    p = engine.ik_params(kinematics=os.environ.get("KIN", "auto"))
    bufs = [(torch.empty((n, 8), device=dev), torch.empty((n, 4), device=dev)) for _ in range(2)]
    f = lambda i=0: engine.ik_solve(tg, neutral, p, out_q8=bufs[i & 1][0], out_aux4=bufs[i & 1][1])
    for _ in range(3): f()
    torch.cuda.synchronize(); ts = []
    for _ in range(7):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    K = 10
    res = {}
    for name, alt in (("same buffers", 0), ("alternating buffers", 1)):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for i in range(K): f(i * alt)
        e1.record(); torch.cuda.synchronize()
        res[name] = e0.elapsed_time(e1) / K
    print(f"ik 2^{lg}: single launch median {statistics.median(ts):.4f} ms min {min(ts):.4f};  {K} back to back: " +
          ", ".join(f"{k} {v:.4f} ms/launch" for k, v in res.items()), flush=True)
    ok = bool((bufs[0][0] == bufs[1][0]).all()) and bool((bufs[0][1] == bufs[1][1]).all())
    print("   both buffer sets identical:", ok, flush=True)
