"""A/B timing of the FP32 IK kernels (development tool, run on the GPU box):
pair (F2) vs lane at 2^24 / 2^22 cold targets, the cfg2 batch through the latency kernel vs the
refill kernel, and the single-call latencies.  PNP_IK_FLUSH_MIN / PNP_IK_OCC are read once per process:
run the script once per setting.  Prints one JSON object."""
import json
import os
import statistics
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from mujoco_panda_pnp_b200 import KinematicData, KinematicModel, KinematicTree, engine, synthetic  # noqa: E402
from mujoco_panda_pnp_b200.envs import FrankaShelfPNPReward  # noqa: E402
from mujoco_panda_pnp_b200.skills import JacobianIKController  # noqa: E402

NEUTRAL = np.array([0.00, 0.41, 0.00, -1.85, 0.00, 2.26, 0.79])


def timed(fn, reps, warm=3):
    for _ in range(warm):
        fn()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    torch.cuda.synchronize()
    for a, b in evs:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    return [a.elapsed_time(b) for a, b in evs]


def main():
    quick = "--quick" in sys.argv
    tree = KinematicTree.from_mjcf()
    engine.set_tree(tree)
    dev = torch.device("cuda")
    neutral = torch.tensor(NEUTRAL, dtype=torch.float32, device=dev)
    out = {"env": {k: os.environ.get(k) for k in ("PNP_IK_FLUSH_MIN", "PNP_IK_OCC", "PNP_IK_SMALL")}}
    peak = max(engine.probe_fp32_peak()[0] for _ in range(2))
    out["fp32_peak_tflops"] = peak
    for log2n in ((24,) if quick else (24, 22, 26)):
        n = 1 << log2n
        targets = torch.empty((n, 3), device=dev)
        for off in range(0, n, 1 << 22):
            m = min(1 << 22, n - off)
            q = synthetic.random_joint_configs(m, tree.lower, tree.upper, seed=1234 + off, device=dev)
            targets[off:off + m] = engine.fk_jac(q, want_quat=False, want_jac=False)[0]
            del q
        q8 = torch.empty((n, 8), device=dev)
        aux = torch.empty((n, 4), device=dev)
        cnt = torch.zeros(4, dtype=torch.int64, device=dev)
        engine.ik_solve(targets, neutral, engine.ik_params(), counters=cnt, out_q8=q8, out_aux4=aux)
        c = cnt.cpu().numpy()
        flop = 500.0 * float(c[3]) + 216.0 * float(c[0])
        for kin in ("spec_pair",) + (() if quick else ("spec_lane",)):
            p = engine.ik_params(kinematics=kin)
            ts = timed(lambda: engine.ik_solve(targets, neutral, p, out_q8=q8, out_aux4=aux), 5 if log2n >= 26 else 10)
            ms = statistics.median(ts)
            out[f"ik_2^{log2n}_{kin}"] = {"ms": ms, "min_ms": min(ts), "gsolves_per_s": float(c[1]) / ms / 1e6,
                                          "tflops": flop / ms / 1e9, "frac": flop / ms / 1e9 / peak}
            if kin != "spec_lane":
                ts = timed(lambda: engine.ik_solve(targets, neutral, p, out_q8=q8, compact=True), 5)
                out[f"ik_2^{log2n}_{kin}"]["compact_ms"] = statistics.median(ts)
        del targets, q8, aux
    # cfg2: 4096 cold targets, one launch
    q = synthetic.random_joint_configs(4096, tree.lower, tree.upper, seed=1234, device=dev)
    t4096 = engine.fk_jac(q, want_quat=False, want_jac=False)[0]
    for kin in ("auto", "spec_lane", "spec_pair"):
        p = engine.ik_params(kinematics=kin)
        ts = timed(lambda: engine.ik_solve(t4096, neutral, p), 50, warm=5)
        out[f"cfg2_4096_{kin}"] = {"us_median": statistics.median(ts) * 1e3, "us_min": min(ts) * 1e3}
    for m in (256, 1024, 8192, 18944):
        tm = synthetic.random_joint_configs(m, tree.lower, tree.upper, seed=7, device=dev)
        tm = engine.fk_jac(tm, want_quat=False, want_jac=False)[0]
        for kin in ("auto", "spec_lane"):
            p = engine.ik_params(kinematics=kin)
            ts = timed(lambda: engine.ik_solve(tm, neutral, p), 20, warm=3)
            out[f"small_{m}_{kin}"] = {"us_median": statistics.median(ts) * 1e3}
    # a CUDA graph of the cfg2 launch (no ticket memset on the small path: a single kernel node)
    g = torch.cuda.CUDAGraph()
    p = engine.ik_params()
    q8 = torch.empty((4096, 8), device=dev)
    aux = torch.empty((4096, 4), device=dev)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        engine.ik_solve(t4096, neutral, p, out_q8=q8, out_aux4=aux)
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            engine.ik_solve(t4096, neutral, p, out_q8=q8, out_aux4=aux)
    ts = timed(g.replay, 50, warm=5)
    out["cfg2_4096_graph"] = {"us_median": statistics.median(ts) * 1e3, "us_min": min(ts) * 1e3}
    # single-call latencies, host to host
    model = KinematicModel.from_xml_path(os.path.join(ROOT, "mujoco_panda_pnp_b200", "assets", "panda_shelf_kinematic.xml"))
    ctl = JacobianIKController(model, KinematicData(model))
    grasp = [np.array(t) for t in [(1.415, 0, 0.73), (1.415, 0, 1.03), (1.415, 0, 0.43)]]
    for t in grasp:
        ctl.solve(t, NEUTRAL)
    t0 = time.perf_counter()
    for _ in range(200):
        for t in grasp:
            ctl.solve(t, NEUTRAL)
    out["solve_one_us"] = (time.perf_counter() - t0) / 600 * 1e6
    env = FrankaShelfPNPReward("dense")
    rows = synthetic.reward_rows(64, seed=0, device="cpu", dtype=torch.float64)
    h = {k: v.numpy() for k, v in rows.items()}
    infos = [dict(ee_pos=h["ee_pos"][i], ee_quat=h["ee_quat"][i], fingers_width=float(h["fingers_width"][i]),
                  task_index=int(h["task_index"][i])) for i in range(64)]
    for i in range(64):
        env.compute_reward(h["achieved_goal"][i], h["desired_goal"][i], infos[i])
    t0 = time.perf_counter()
    for _ in range(50):
        for i in range(64):
            env.compute_reward(h["achieved_goal"][i], h["desired_goal"][i], infos[i])
    out["reward_one_us"] = (time.perf_counter() - t0) / 3200 * 1e6
    import ctypes

    from mujoco_panda_pnp_b200 import _lib
    lib = _lib.load()
    rp = engine.reward_params()
    r, sc, bits = ctypes.c_float(), ctypes.c_float(), ctypes.c_uint32()
    ag, dg, ee, eq = (np.ascontiguousarray(h[k][0]) for k in ("achieved_goal", "desired_goal", "ee_pos", "ee_quat"))
    ctx = engine.host_ctx(0)
    t0 = time.perf_counter()
    for _ in range(2000):
        lib.pnp_reward_one_host_f64(ctx, ag.ctypes.data, dg.ctypes.data, ee.ctypes.data, eq.ctypes.data, 0.03, 1,
                                    ctypes.byref(rp), ctypes.byref(r), ctypes.byref(sc), ctypes.byref(bits))
    out["reward_one_c_call_us"] = (time.perf_counter() - t0) / 2000 * 1e6
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
