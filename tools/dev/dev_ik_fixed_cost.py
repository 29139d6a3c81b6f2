"""Where does the fixed ~0.15 ms per launch of the pair kernel come from?  (development tool, GPU box)
t(n) is fitted as a + b*n over n = 2^21 .. 2^25 for three workloads:
  cold/100   the bench workload (0.2 % of the queries run 100 passes)
  cold/30    same targets, max_iters = 30 (the longest query is 30 passes: a 3x shorter drain)
  easy       targets a few cm from FK(neutral): every query takes 3-6 passes (no drain to speak of)
If `a` shrinks with the length of the longest query, it is the drain; what stays is launch ramp + memset."""
import json
import os
import statistics
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from mujoco_panda_pnp_b200 import KinematicTree, engine, synthetic  # noqa: E402

NEUTRAL = [0.00, 0.41, 0.00, -1.85, 0.00, 2.26, 0.79]


def timed(fn, reps=7):
    for _ in range(2):
        fn()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    torch.cuda.synchronize()
    for a, b in evs:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    return statistics.median(a.elapsed_time(b) for a, b in evs)


def main():
    tree = KinematicTree.from_mjcf()
    engine.set_tree(tree)
    dev = torch.device("cuda")
    neutral = torch.tensor(NEUTRAL, dtype=torch.float32, device=dev)
    nmax = 1 << 25
    cold = torch.empty((nmax, 3), device=dev)
    for off in range(0, nmax, 1 << 22):
        q = synthetic.random_joint_configs(1 << 22, tree.lower, tree.upper, seed=1234 + off, device=dev)
        cold[off:off + (1 << 22)] = engine.fk_jac(q, want_quat=False, want_jac=False)[0]
    g = torch.Generator(device=dev); g.manual_seed(1)
    qe = neutral + 0.15 * (torch.rand((nmax, 7), generator=g, device=dev) - 0.5)
    easy = engine.fk_jac(qe.contiguous(), want_quat=False, want_jac=False)[0]
    del qe
    q8 = torch.empty((nmax, 8), device=dev); aux = torch.empty((nmax, 4), device=dev)
    out = {}
    for name, tg, mi in (("cold/100", cold, 100), ("cold/30", cold, 30), ("easy", easy, 100)):
        p = engine.ik_params(max_iters=mi, kinematics="spec_pair")
        xs, ys, its = [], [], None
        for lg in (21, 22, 23, 24, 25):
            n = 1 << lg
            cnt = torch.zeros(4, dtype=torch.int64, device=dev)
            engine.ik_solve(tg[:n], neutral, p, counters=cnt, out_q8=q8[:n], out_aux4=aux[:n])
            its = float(cnt[3]) / n
            ms = timed(lambda: engine.ik_solve(tg[:n], neutral, p, out_q8=q8[:n], out_aux4=aux[:n]))
            xs.append(n); ys.append(ms)
        b, a = np.polyfit(np.array(xs, float), np.array(ys), 1)
        out[name] = {"ms": dict(zip(map(str, xs), [round(y, 4) for y in ys])), "fixed_ms": round(float(a), 4),
                     "ns_per_query": round(float(b) * 1e6, 4), "mean_iterations": round(its, 2)}
    # an empty-ish launch: 2^21 queries that converge on their first pass (targets = FK(neutral))
    t0 = engine.fk_jac(neutral[None].expand(1 << 21, 7).contiguous(), want_quat=False, want_jac=False)[0]
    p = engine.ik_params(kinematics="spec_pair")
    out["first_pass_2^21"] = {"ms": timed(lambda: engine.ik_solve(t0, neutral, p, out_q8=q8[:1 << 21], out_aux4=aux[:1 << 21]))}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
