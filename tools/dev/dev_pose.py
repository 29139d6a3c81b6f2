"""Pose-mode IK timing (development tool, GPU box).  Prints one JSON object.
(Round 2 measured a two-queries-per-lane packed variant of this kernel with this script: 230 registers, 2-3 blocks per
SM, 1.067 ms against 1.091 ms for 2^22 poses - the per-query parts (mju_mat2Quat's case selection, atan2, the serial
6x6 LDL^T) dominate and do not pack; dropped.)"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from mujoco_panda_pnp_b200 import KinematicTree, engine, synthetic  # noqa: E402

tree = KinematicTree.from_mjcf()
engine.set_tree(tree)
dev = torch.device("cuda")
neutral = torch.tensor(synthetic.NEUTRAL_Q, device=dev)
peak = max(engine.probe_fp32_peak()[0] for _ in range(2))
out = {"fp32_peak_tflops": peak}
for lg in (20, 22):
    n = 1 << lg
    qp = synthetic.reachable_move_envs(n, tree.lower, tree.upper, seed=5, device=dev, spread=0.5)["q_goal"]
    ppos, pquat, _ = engine.fk_jac(qp, want_jac=False)
    for kin in ("spec_lane",):
        p = engine.ik_params(kinematics=kin)
        cnt = torch.zeros(4, dtype=torch.int64, device=dev)
        engine.ik_pose_solve(ppos, pquat, neutral, p, counters=cnt)
        torch.cuda.synchronize()
        c = cnt.cpu().numpy()
        for _ in range(2):
            engine.ik_pose_solve(ppos, pquat, neutral, p)
        ts = []
        for _ in range(7):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); engine.ik_pose_solve(ppos, pquat, neutral, p); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[len(ts) // 2]
        tf = 950.0 * float(c[3]) / ms / 1e9
        out[f"pose_2^{lg}_{kin}"] = {"ms": ms, "g_solves_per_s": n / ms / 1e6, "converged": float(c[1]) / n,
                                     "mean_iterations": float(c[3]) / n, "tflops_algorithmic_950": tf, "frac": tf / peak,
                                     "note": "includes the torch.empty of 7 output arrays"}
print(json.dumps(out, indent=1))
