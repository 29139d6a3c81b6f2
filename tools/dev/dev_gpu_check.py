"""Developer scratch check on a GPU box: smoke + rough timings (not the bench)."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch
import __graft_entry__ as g
g.smoke()
from mujoco_panda_pnp_b200 import engine, synthetic, KinematicTree
from mujoco_panda_pnp_b200.tree import DEFAULT_ASSET

tree = KinematicTree.from_mjcf()
print("specialized:", engine.set_tree(tree))
print("fp32 peak probe:", engine.probe_fp32_peak())
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

for logn in (12, 16, 20, 22, 24):
    n = 1 << logn
    qs = synthetic.random_joint_configs(n, tree.lower, tree.upper, seed=1234, device=dev)
    tg, _, _ = engine.fk_jac(qs, want_quat=False, want_jac=False)
    for kin in ("specialized", "generic"):
        p = engine.ik_params(kinematics=kin)
        cnt = torch.zeros(4, dtype=torch.int64, device=dev)
        r = engine.ik_solve(tg, neutral, p, counters=cnt)
        torch.cuda.synchronize()
        c = cnt.cpu().numpy()
        best, avg = timeit(lambda: engine.ik_solve(tg, neutral, p))
        flop = 500.0 * c[3] + 216.0 * c[0]
        print(f"ik f32 {kin:11s} n=2^{logn}: best {best:8.3f} ms avg {avg:8.3f} ms -> {n / best / 1e3:9.2f} M solves/s  "
              f"conv {c[1] / c[0]:.4f} mean it {c[3] / c[0]:.2f}  alg {flop / best / 1e9:.2f} TFLOP/s")
    if logn <= 20:
        p = engine.ik_params(kinematics="specialized")
        best, avg = timeit(lambda: engine.ik_solve(tg.double(), neutral.double(), p))
        print(f"ik f64 spec n=2^{logn}: best {best:8.3f} ms -> {n / best / 1e3:9.2f} M solves/s")

for logn in (20, 24):
    n = 1 << logn
    for dt in (torch.float32, torch.float64):
        rows = synthetic.reward_rows(n, seed=0, device=dev, dtype=dt)
        for rt in ("dense", "sparse"):
            p = engine.reward_params(rt)
            args = [rows[k] for k in ("achieved_goal", "desired_goal", "ee_pos", "ee_quat", "fingers_width", "task_index")]
            out = torch.empty(n, dtype=torch.float32, device=dev)
            best, avg = timeit(lambda: engine.reward(*args, p, want_success=False, out=out))
            bpr = (60 if dt == torch.float32 else 116) + 4
            print(f"reward {str(dt):14s} {rt:6s} n=2^{logn}: best {best:7.3f} ms avg {avg:7.3f} -> {n / best / 1e6:8.2f} G rows/s {n * bpr / best / 1e6:8.1f} GB/s")

# waypoints
n = 1 << 20
w = synthetic.waypoint_envs(n, seed=0, device=dev)
p = engine.ik_params()
cnt = torch.zeros(4, dtype=torch.int64, device=dev)
r = engine.ik_waypoints(w["q_start"], w["goal"], 50, p, counters=cnt); torch.cuda.synchronize()
c = cnt.cpu().numpy()
best, avg = timeit(lambda: engine.ik_waypoints(w["q_start"], w["goal"], 50, p), warm=1, rep=3)
print(f"waypoints n=2^20 x50: best {best:.2f} ms -> {c[0] / best / 1e3:.1f} M warm solves/s; mean it {c[3] / c[0]:.2f} conv {c[1] / c[0]:.4f} accepted mean {r['n_accepted'].float().mean().item():.2f}")

# host path
n = 1 << 22
tgh = tg[:n].cpu().pin_memory() if tg.shape[0] >= n else None
qs = synthetic.random_joint_configs(n, tree.lower, tree.upper, seed=1234, device=dev)
tgd, _, _ = engine.fk_jac(qs, want_quat=False, want_jac=False)
tgh = tgd.cpu().pin_memory()
outs = dict(q8=torch.empty((n, 8), dtype=torch.float32).pin_memory().numpy(), aux4=torch.empty((n, 4)).pin_memory().numpy())
p = engine.ik_params()
for _ in range(2): engine.ik_solve_host(tgh, np.array(synthetic.NEUTRAL_Q, dtype=np.float32), p, out=outs)
t0 = time.perf_counter(); r = engine.ik_solve_host(tgh, np.array(synthetic.NEUTRAL_Q, dtype=np.float32), p, out=outs); t1 = time.perf_counter()
print(f"ik host e2e n=2^22: {(t1 - t0) * 1e3:.2f} ms -> {n / (t1 - t0) / 1e6:.1f} M solves/s, counters {r['counters']}")
