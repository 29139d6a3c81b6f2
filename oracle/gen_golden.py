"""ORACLE (test infrastructure): generate tests/golden/* from the reference's own Python.

Run in the build container only (needs /root/reference):

    python -m oracle.gen_golden

What is executed to produce the expected outputs:

* IK      -> /root/reference/panda_mujoco_gym/skills/ik_solver.py  JacobianIKController.solve
* reward  -> /root/reference/panda_mujoco_gym/envs/panda_env.py    FrankaEnv.compute_reward,
             FrankaEnv._is_success, FrankaEnv.goal_distance, class constants *_QUAT

both imported unmodified through oracle/ref_harness.py (third-party natives stubbed, see there).
FK / Jacobian vectors come from oracle/mj_oracle.py (no reference-side Python exists for
them: they are MuJoCo C routines); the one real-MuJoCo value, home_wpt, is stored beside them.
The model is compiled from the reference's assets/shelf_pnp.xml.
"""

from __future__ import annotations

import json
import os
import pickle
import sys

import numpy as np

from . import ik_oracle, mj_oracle, ref_harness

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden")
NEUTRAL = np.array([0.00, 0.41, 0.00, -1.85, 0.00, 2.26, 0.79])
HOME_WPT = np.array([1.23843967, 0.0, 0.49740014])  # scripts/execute_pnp.py:38 (real MuJoCo)


def gen_fk(model):
    rng = np.random.default_rng(7)
    lo, hi = model.jnt_range[:7, 0], model.jnt_range[:7, 1]
    qs = [NEUTRAL, np.zeros(7), lo.copy(), hi.copy(), 0.5 * (lo + hi)]
    qs += list(rng.uniform(lo, hi, size=(59, 7)))
    qs = np.array(qs)
    data = mj_oracle.MjData(model)
    pos, mat, quat, jac = [], [], [], []
    for q in qs:
        p, m, qu, j = ik_oracle.fk_site(model, data, q)
        pos.append(p), mat.append(m), quat.append(qu), jac.append(j)
    np.savez(
        os.path.join(OUT, "fk_jac_golden.npz"),
        q=qs, pos=np.array(pos), mat=np.array(mat), quat=np.array(quat), jac=np.array(jac),
        home_wpt=HOME_WPT, neutral=NEUTRAL,
    )
    print("fk_jac_golden: %d poses; |FK(neutral)-home_wpt| = %.2e" % (len(qs), np.abs(pos[0] - HOME_WPT).max()))


def gen_ik(model):
    ref_ik = ref_harness.reference_ik_module()
    ctl = ref_ik.JacobianIKController(model, mj_oracle.MjData(model))
    data = mj_oracle.MjData(model)
    lo, hi = model.jnt_range[:7, 0], model.jnt_range[:7, 1]
    default = dict(max_iters=100, pos_thresh=1e-3, damping=1e-2, step_limit=0.1)
    cases = []

    def add(tag, target, q_init, **kw):
        cases.append((tag, np.asarray(target, float), np.asarray(q_init, float), {**default, **kw}))

    # cfg1: grasp poses = cube body pos + [0.015, 0, 0] (execute_pnp.py:33, shelf_pnp.xml:61,67,73)
    for k, t in enumerate([(1.415, 0, 0.73), (1.415, 0, 1.03), (1.415, 0, 0.43)]):
        add(f"cfg1_cube{k + 1}", t, NEUTRAL)
    add("cfg1_home", HOME_WPT, NEUTRAL)
    # test/ik_test.py:26-38 variant
    add("ik_test", HOME_WPT + np.array([0.1, 0, 0]), NEUTRAL, pos_thresh=1e-4, damping=0.05)
    # cfg2: reachable random targets, cold start from neutral
    rng = np.random.default_rng(0)
    qstar = rng.uniform(lo, hi, size=(4096, 7))
    for k in range(96):
        add(f"cfg2_{k}", ik_oracle.fk_site(model, data, qstar[k])[0], NEUTRAL)
    # warm starts (MoveIKSkill step: <= 1 cm away, move.py:114-128)
    rng2 = np.random.default_rng(1)
    for k in range(16):
        q0 = np.clip(NEUTRAL + rng2.uniform(-0.3, 0.3, 7), lo, hi)
        p0 = ik_oracle.fk_site(model, data, q0)[0]
        d = rng2.normal(size=3)
        add(f"warm_{k}", p0 + 0.01 * d / np.linalg.norm(d), q0)
    # loop exhaustion / unreachable / different parameters
    add("unreachable", (2.5, 0.0, 0.5), NEUTRAL)
    add("unreachable_low", (0.6, 0.0, -0.5), NEUTRAL)
    add("max_iters_3", (1.415, 0, 0.73), NEUTRAL, max_iters=3)
    add("max_iters_7_exact", (1.415, 0, 0.73), NEUTRAL, max_iters=7)
    add("max_iters_6_short", (1.415, 0, 0.73), NEUTRAL, max_iters=6)
    add("big_step", (1.415, 0, 1.03), NEUTRAL, step_limit=0.5)
    add("tight", (1.3, 0.2, 0.6), NEUTRAL, pos_thresh=1e-5, damping=1e-3, max_iters=200)
    add("q_init_outside_limits", (1.3, -0.1, 0.7), hi + 0.2)

    keys = ("success", "q", "final_pos", "pos_error", "iterations", "converged")
    out = {k: [] for k in keys}
    for tag, target, q_init, kw in cases:
        r = ctl.solve(target, q_init, **kw)
        for k in keys:
            out[k].append(getattr(r, k))
    np.savez(
        os.path.join(OUT, "ik_reference_golden.npz"),
        tag=np.array([c[0] for c in cases]),
        target=np.array([c[1] for c in cases]),
        q_init=np.array([c[2] for c in cases]),
        max_iters=np.array([c[3]["max_iters"] for c in cases], dtype=np.int32),
        pos_thresh=np.array([c[3]["pos_thresh"] for c in cases]),
        damping=np.array([c[3]["damping"] for c in cases]),
        step_limit=np.array([c[3]["step_limit"] for c in cases]),
        success=np.array(out["success"], dtype=bool),
        converged=np.array(out["converged"], dtype=bool),
        q=np.array(out["q"]),
        final_pos=np.array(out["final_pos"]),
        pos_error=np.array(out["pos_error"], dtype=np.float64),
        iterations=np.array(out["iterations"], dtype=np.int32),
    )
    it = np.array(out["iterations"])
    print("ik_reference_golden: %d cases, converged %d, iterations mean %.2f max %d"
          % (len(cases), int(np.sum(out["converged"])), it.mean(), it.max()))


def gen_reward():
    sys.path.insert(0, ROOT)
    from mujoco_panda_pnp_b200 import synthetic  # input generator only
    import torch

    env_mod = ref_harness.reference_env_module()
    rows = synthetic.reward_rows(4096, seed=0, device="cpu", dtype=torch.float64, n_adversarial=512)
    rows = {k: v.numpy() for k, v in rows.items()}
    # SURVEY.md App. C known-answer rows first (row 0 reproduces the pickle's old_reward)
    H = env_mod.FrankaEnv.HORIZONTAL_QUAT
    kat = [
        ((1.46172, 0.17519, 0.01989), (1, -0.1, 0.3), (1.3847, 0.35938, 0.58267), 0.07983, (0, 1, 0, 0), 0),
        ((1.4, 0, 0.73), (1, -0.1, 0.3), (1.4, 0.02, 0.73), 0.08, (0, 1, 0, 0), 0),
        ((1.4, 0, 0.73), (1, -0.1, 0.3), (1.41, 0, 0.73), 0.04, tuple(H), 0),
        ((1.0, 0.1, 0.32), (1, -0.1, 0.3), (1.0, 0.1, 0.33), 0.04, (1, 0, 0, 0), 1),
        ((1.0, -0.1, 0.32), (1, -0.1, 0.3), (1.0, -0.1, 0.33), 0.04, (0, 1, 0, 0), 2),
        ((1.0, -0.1, 0.32), (1, -0.1, 0.3), (1.0, -0.1, 0.45), 0.08, (0, 1, 0, 0), 0),
    ]
    for i, (ag, dg, ee, w, q, t) in enumerate(kat):
        rows["achieved_goal"][i] = ag
        rows["desired_goal"][i] = dg
        rows["ee_pos"][i] = ee
        rows["fingers_width"][i] = w
        rows["ee_quat"][i] = q
        rows["task_index"][i] = t
    n = len(rows["task_index"])
    out = {}
    for rt in ("dense", "sparse"):
        probe = ref_harness.RewardProbe(env_mod, reward_type=rt)
        r = np.empty(n, dtype=np.float32)
        s = np.empty(n, dtype=np.float32)
        for i in range(n):
            r[i] = probe.reward(rows["achieved_goal"][i], rows["desired_goal"][i], rows["ee_pos"][i],
                                rows["ee_quat"][i], rows["fingers_width"][i], rows["task_index"][i])
            s[i] = probe.success(rows["achieved_goal"][i], rows["desired_goal"][i])
        out[f"reward_{rt}"] = r
        out["is_success"] = s
    gd = env_mod.FrankaEnv.goal_distance(None, rows["achieved_goal"], rows["desired_goal"])
    np.savez(
        os.path.join(OUT, "reward_reference_golden.npz"),
        **rows, **out, goal_distance=gd,
        horizontal_quat=env_mod.FrankaEnv.HORIZONTAL_QUAT, vertical_quat=env_mod.FrankaEnv.VERTICAL_QUAT,
        n_tasks=3, initial_object_height=0.001, distance_threshold=0.05, high_pick_z=0.35,
    )
    d = out["reward_dense"]
    print("reward_reference_golden: %d rows; dense in [%.3f, %.3f]; row0 bits %s; placed %d; "
          "gripped-ish (>1.5) %d; lifted-ish (>5.5) %d; sparse -0.0 count %d"
          % (n, d.min(), d.max(), hex(d[:1].view(np.uint32)[0]), int(out["is_success"].sum()),
             int((d > 1.5).sum()), int((d > 5.5).sum()),
             int((out["reward_sparse"].view(np.uint32) == 0x80000000).sum())))


def random_obs_states(n, model, seed=11):
    """Kinematic states for _get_obs: arm inside its limits, joint velocities, open/closed fingers,
    a free cube with arbitrary pose (incl. near-gimbal orientations) and twist."""
    rng = np.random.default_rng(seed)
    lo, hi = model.jnt_range[:7, 0], model.jnt_range[:7, 1]
    st = dict(
        q_arm=rng.uniform(lo, hi, (n, 7)), qvel_arm=rng.normal(0, 0.5, (n, 7)),
        fingers=rng.uniform(0, 0.04, (n, 2)), obj_pos=rng.normal([1.3, 0, 0.6], 0.3, (n, 3)),
        obj_quat=rng.normal(size=(n, 4)), obj_vel=rng.normal(0, 0.3, (n, 6)),
        goal=np.array([[1.0, -0.1, 0.3], [1.0, 0.0, 0.3], [1.0, 0.1, 0.3]])[rng.integers(0, 3, n)],
        obj_index=rng.integers(0, 3, n),
    )
    st["obj_quat"] /= np.linalg.norm(st["obj_quat"], axis=1, keepdims=True)
    # pitch = +-90 deg rows exercise mat2euler's degenerate branch (cy ~ 0)
    s = np.sqrt(0.5)
    st["obj_quat"][:4] = [[s, 0, s, 0], [s, 0, -s, 0], [1, 0, 0, 0], [0, 1, 0, 0]]
    st["qvel_arm"][4] = 0.0
    return st


def gen_obs(model):
    from . import obs_oracle

    env_mod = ref_harness.reference_env_module()
    data = mj_oracle.MjData(model)
    probe = ref_harness.ObsProbe(env_mod, model, data)
    n = 256
    st = random_obs_states(n, model)
    obs, ag, dg = np.empty((n, 19)), np.empty((n, 3)), np.empty((n, 3))
    for i in range(n):
        name = f"cube{int(st['obj_index'][i]) + 1}"
        obs_oracle.set_state(model, data, st["q_arm"][i], st["qvel_arm"][i], st["fingers"][i], name,
                             st["obj_pos"][i], st["obj_quat"][i], st["obj_vel"][i])
        o = probe.get_obs(name, st["goal"][i].copy())
        obs[i], ag[i], dg[i] = o["observation"], o["achieved_goal"], o["desired_goal"]
    np.savez(os.path.join(OUT, "obs_reference_golden.npz"), **st, observation=obs, achieved_goal=ag,
             desired_goal=dg, dt=0.05)
    print("obs_reference_golden: %d states; |obs| max %.3f" % (n, np.abs(obs).max()))


def move_cases(model):
    """Planner cases (q_start, target): ordinary reachable moves of different lengths, the BT demo
    waypoints, a start already within 1 cm, and moves towards the workspace boundary whose IK
    solves fail part of the way (step halving, double failure increment, fallback strategy 1) but
    that still terminate because point_count reaches max_traj_points.  Unreachable targets are NOT
    included: the reference loop never terminates on them (see oracle/c/pnp_oracle.c)."""
    lo, hi = model.jnt_range[:7, 0], model.jnt_range[:7, 1]
    rng = np.random.default_rng(3)
    n = 3000
    tg = rng.uniform([0.3, -0.8, -0.1], [1.75, 0.8, 1.5], (n, 3))
    q0 = np.clip(NEUTRAL + rng.uniform(-0.4, 0.4, (n, 7)), lo, hi)
    cases = [(NEUTRAL, np.array(t, float)) for t in
             [(1.415, 0, 0.73), (1.415, 0, 1.03), (1.415, 0, 0.43), (1.2, 0.0, 0.78), (1.24, 0.0, 0.5), (1.0, -0.1, 0.36)]]
    for i in (0, 1, 2, 3, 4, 10, 12, 19):        # ordinary random moves (all terminate)
        cases.append((q0[i], tg[i]))
    for i in (991, 592, 1184, 1026):             # failing solves on the way, terminate at 200 points
        cases.append((q0[i], tg[i]))
    return cases


def gen_move(model):
    import signal

    def on_alarm(signum, frame):
        raise TimeoutError("reference MoveIKSkill.reset did not terminate")

    signal.signal(signal.SIGALRM, on_alarm)
    cases = move_cases(model)
    cap = 208
    traj = np.zeros((len(cases), cap, 3))
    lens = np.zeros(len(cases), dtype=np.int32)
    for k, (q_start, target) in enumerate(cases):
        signal.alarm(300)
        t = ref_harness.reference_move_plan(model, q_start, target)
        signal.alarm(0)
        lens[k] = len(t)
        traj[k, : len(t)] = t
    np.savez(os.path.join(OUT, "move_reference_golden.npz"), q_start=np.array([c[0] for c in cases]),
             target=np.array([c[1] for c in cases]), traj=traj, traj_len=lens,
             pos_thresh=0.01, max_traj_points=200, step_size=0.01)
    print("move_reference_golden: %d cases, trajectory lengths %s" % (len(cases), lens.tolist()))


def gen_vecnormalize():
    class _Stub:
        def __init__(self, *a, **k):
            pass

        def __setstate__(self, s):
            self.__dict__.update(s if isinstance(s, dict) else {"state": s})

    class _U(pickle.Unpickler):
        def find_class(self, mod, name):
            try:
                return super().find_class(mod, name)
            except Exception:
                return type(name, (_Stub,), {"__module__": mod})

    path = os.path.join(ref_harness.REFERENCE_ROOT, "scripts", "checkpoints", "tqc_dense_vecnormalize_200000_steps.pkl")
    with open(path, "rb") as fh:
        vn = _U(fh).load()
    stats = {
        "source": "scripts/checkpoints/tqc_dense_vecnormalize_200000_steps.pkl",
        "old_reward": [float(x) for x in vn.old_reward],
        "old_reward_bits": [hex(int(np.float32(x).view(np.uint32))) for x in vn.old_reward],
        "clip_obs": float(vn.clip_obs), "epsilon": float(vn.epsilon), "gamma": float(vn.gamma),
        "obs_rms": {k: {"mean": v.mean.tolist(), "var": v.var.tolist(), "count": float(v.count)}
                    for k, v in vn.obs_rms.items()},
    }
    with open(os.path.join(OUT, "vecnormalize_stats.json"), "w") as fh:
        json.dump(stats, fh, indent=1)
    print("vecnormalize_stats: old_reward", stats["old_reward"], stats["old_reward_bits"][0])


def main():
    if not ref_harness.available():
        raise SystemExit("needs /root/reference (build container only)")
    os.makedirs(OUT, exist_ok=True)
    model = mj_oracle.MjModel.from_xml_path(ref_harness.reference_xml_path())
    gen_fk(model)
    gen_ik(model)
    gen_reward()
    gen_obs(model)
    gen_vecnormalize()
    gen_move(model)


if __name__ == "__main__":
    main()
