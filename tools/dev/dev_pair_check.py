"""Developer scratch check: pair (FFMA2) IK kernel vs lane kernels - bit identity and timing."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch
from mujoco_panda_pnp_b200 import engine, synthetic, KinematicTree

tree = KinematicTree.from_mjcf()
print("specialized:", engine.set_tree(tree))
print("fp32 peak probe:", engine.probe_fp32_peak())
dev = torch.device("cuda")
neutral = torch.tensor(synthetic.NEUTRAL_Q, device=dev)

def timeit(fn, warm=2, rep=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(rep):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), sum(ts) / len(ts)

logs = [int(x) for x in (sys.argv[1:] or ["12", "16", "18", "20", "22", "24"])]
for logn in logs:
    n = (1 << logn) - (3 if logn < 20 else 0)
    qs = synthetic.random_joint_configs(n, tree.lower, tree.upper, seed=1234, device=dev)
    tg, _, _ = engine.fk_jac(qs, want_quat=False, want_jac=False)
    res = {}
    for kin in ("generic", "spec_lane", "spec_pair"):
        p = engine.ik_params(kinematics=kin)
        cnt = torch.zeros(4, dtype=torch.int64, device=dev)
        r = engine.ik_solve(tg, neutral, p, counters=cnt)
        torch.cuda.synchronize()
        res[kin] = r
        c = cnt.cpu().numpy()
        best, avg = timeit(lambda: engine.ik_solve(tg, neutral, p))
        flop = 500.0 * c[3] + 216.0 * c[0]
        print(f"ik f32 {kin:11s} n={n}: best {best:8.3f} ms avg {avg:8.3f} ms -> {n / best / 1e3:9.2f} M solves/s  "
              f"conv {c[1] / c[0]:.4f} mean it {c[3] / c[0]:.3f} N {c[0]}  alg {flop / best / 1e9:.2f} TFLOP/s")
    a, b, l = res["spec_pair"], res["spec_lane"], res["generic"]
    for k in ("q", "final_pos", "pos_error", "iterations", "converged", "success"):
        x, y, z = getattr(a, k), getattr(b, k), getattr(l, k)
        same = bool((x == y).all().item())
        close = (x.double() - z.double()).abs().max().item()
        print(f"   {k:10s} pair==lane bitwise: {same}   max|pair-legacy| {close:.3e}")
    # per-query q_init + unpacked outputs
    if logn <= 20:
        qi = (neutral[None, :] + 0.05 * torch.randn(n, 7, device=dev)).contiguous()
        ra = engine.ik_solve(tg, qi, engine.ik_params(kinematics="spec_pair"))
        rb = engine.ik_solve(tg, qi, engine.ik_params(kinematics="spec_lane"))
        print("   per-query q_init pair==lane:", bool((ra.q == rb.q).all().item()), bool((ra.iterations == rb.iterations).all().item()))
