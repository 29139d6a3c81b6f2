"""Developer scratch: stress the overlapping-launch machinery (programmatic dependent launches, alternating tickets,
drain hand-over).  Random sequences of IK solves of different sizes / kernels / output layouts on two streams, nothing
synchronised in between; every result is compared with the same solve run alone.  Prints the number of mismatching launches."""
import os, sys, random
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from mujoco_panda_pnp_b200 import KinematicTree, engine, synthetic


def run(seed=0, rounds=6, launches=24):
    tree = KinematicTree.from_mjcf(); engine.set_tree(tree)
    dev = torch.device("cuda")
    neutral = torch.tensor(synthetic.NEUTRAL_Q, dtype=torch.float32, device=dev)
    sizes = [4096, 30_000, 60_000, 300_000, (1 << 20) + 3, 3_000_000]
    rng = random.Random(seed)
    inputs = {}
    for n in sizes:
        q = synthetic.random_joint_configs(n, tree.lower, tree.upper, seed=n % 977, device=dev)
        t = engine.fk_jac(q, want_quat=False, want_jac=False)[0]
        t[::173] = torch.tensor([2.5, 0.0, 0.5], device=dev)
        qi = (neutral + 0.1 * torch.randn((n, 7), device=dev)).contiguous()
        inputs[n] = (t, qi)
    ref = {}
    def reference(n, per_query, kin):
        key = (n, per_query)
        if key not in ref:
            t, qi = inputs[n]
            r = engine.ik_solve(t, qi if per_query else neutral, engine.ik_params(kinematics="spec_lane" if n > 20000 else "auto"))
            torch.cuda.synchronize()
            ref[key] = (r.q.clone(), r.iterations.clone(), r.final_pos.clone())
        return ref[key]
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    bad = total = 0
    for rnd in range(rounds):
        plan = []
        for _ in range(launches):
            n = rng.choice(sizes); per_query = rng.random() < 0.3
            kin = rng.choice(["auto", "auto", "spec_pair", "spec_lane"]) if n > 40000 else "auto"
            plan.append((n, per_query, kin, rng.randrange(2), rng.random() < 0.15))
        for n, pq, kin, _, _ in plan:
            reference(n, pq, kin)
        torch.cuda.synchronize()
        got = []
        for n, pq, kin, si, reuse in plan:
            t, qi = inputs[n]
            with torch.cuda.stream(streams[si]):
                kw = {}
                if reuse and got and got[-1][0] == n and got[-1][4] == si:  # write into the previous launch's buffers: must serialise
                    kw = dict(out_q8=got[-1][3][0], out_aux4=got[-1][3][1])
                    got[-1] = None
                else:
                    kw = dict(out_q8=torch.empty((n, 8), device=dev), out_aux4=torch.empty((n, 4), device=dev))
                r = engine.ik_solve(t, qi if pq else neutral, engine.ik_params(kinematics=kin), **kw)
                got.append((n, pq, r, (kw["out_q8"], kw["out_aux4"]), si))
        torch.cuda.synchronize()
        for g in got:
            if g is None:
                continue
            n, pq, r, _, _ = g
            q, it, fp = ref[(n, pq)]
            total += 1
            if not (torch.equal(r.q, q) and torch.equal(r.iterations, it) and torch.equal(r.final_pos, fp)):
                bad += 1
                print("MISMATCH", n, pq, int((r.iterations != it).sum()), flush=True)
    print(f"stress: {total} launches checked, {bad} mismatches")
    return total, bad


if __name__ == "__main__":
    t, b = run(int(sys.argv[1]) if len(sys.argv) > 1 else 0, int(sys.argv[2]) if len(sys.argv) > 2 else 6)
    sys.exit(1 if b else 0)
