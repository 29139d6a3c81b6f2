"""GPU parity: device-side MoveIKSkill.reset planner (skills/move.py:76-191) vs the trajectories
of the reference's own code (tests/golden/move_reference_golden.npz) and the C oracle."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, NEUTRAL
from mujoco_panda_pnp_b200 import KinematicTree, engine, synthetic
from oracle import c_oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def golden(cuda_lib):
    engine.set_tree(KinematicTree.from_mjcf())
    return np.load(os.path.join(GOLDEN, "move_reference_golden.npz"))


@pytest.mark.parametrize("kin", ["specialized", "generic"])
def test_fp64_planner_reproduces_reference_trajectories(golden, kin):
    g = golden
    out = engine.move_ik_plan(torch.tensor(g["q_start"], device="cuda"), torch.tensor(g["target"], device="cuda"),
                              engine.ik_params(kinematics=kin), traj_cap=g["traj"].shape[1])
    np.testing.assert_array_equal(out["traj_len"].cpu().numpy(), g["traj_len"])
    assert int(out["status"].abs().sum()) == 0
    traj = out["traj"].cpu().numpy()
    for k in range(len(g["traj_len"])):
        L = int(g["traj_len"][k])
        np.testing.assert_allclose(traj[k, :L], g["traj"][k, :L], atol=1e-8)


def test_fp32_planner_follows_reference_trajectories(golden, oracle_chain):
    """FP32 product path: same waypoint counts on the ordinary moves, every waypoint within 1e-4 m
    (north_star EE tolerance); moves with failing solves may branch differently and are compared
    on their structure (length, monotone approach) only."""
    g = golden
    out = engine.move_ik_plan(torch.tensor(g["q_start"], dtype=torch.float32, device="cuda"),
                              torch.tensor(g["target"], dtype=torch.float32, device="cuda"), engine.ik_params(),
                              traj_cap=g["traj"].shape[1])
    tl, traj = out["traj_len"].cpu().numpy(), out["traj"].double().cpu().numpy()
    ref = c_oracle.move_plan(oracle_chain, g["q_start"], g["target"], traj_cap=g["traj"].shape[1])
    clean = ref["n_solves"] < ref["traj_len"]  # every solve accepted (points = start + solves [+ final target])
    assert clean.sum() >= 12 and (~clean).sum() >= 4
    for k in np.nonzero(clean)[0]:
        L = int(g["traj_len"][k])
        assert tl[k] == L
        assert np.abs(traj[k, :L] - g["traj"][k, :L]).max() < 1e-4
    for k in np.nonzero(~clean)[0]:
        assert tl[k] == 202 and int(out["status"][k]) == 0
    # q_final really is the joint solution of the last accepted waypoint
    qf = out["q_final"].double().cpu().numpy()
    ee = c_oracle.fk_jac(oracle_chain, qf)[0]
    for k in range(len(tl)):
        # the last accepted waypoint is the final point, or the one before an appended target
        cands = [traj[k, tl[k] - 1]] + ([traj[k, tl[k] - 2]] if tl[k] > 1 else [])
        assert min(np.linalg.norm(ee[k] - c) for c in cands) < 2e-5


def test_planner_bounds_unreachable_targets_and_overflow(golden, oracle_chain):
    q0 = torch.tensor(np.tile(NEUTRAL, (3, 1)), device="cuda")
    tg = torch.tensor([[2.0, 0.0, 0.5], [1.415, 0.0, 0.73], [1.24, 0.0, 0.5]], dtype=torch.float64, device="cuda")
    out = engine.move_ik_plan(q0, tg, engine.ik_params(), max_outer=300, traj_cap=64)
    st, tl = out["status"].tolist(), out["traj_len"].tolist()
    assert st[0] & 2 and st[0] & 4 and tl[0] > 64       # capped + overflow, like the bounded oracle
    assert st[1] == 0 and tl[1] == 45 and st[2] == 0 and tl[2] == 1
    ref = c_oracle.move_plan(oracle_chain, q0.cpu().numpy(), tg.cpu().numpy(), max_outer=300, traj_cap=64)
    np.testing.assert_array_equal(out["traj_len"].cpu().numpy(), ref["traj_len"])
    np.testing.assert_array_equal(out["n_solves"].cpu().numpy(), ref["n_solves"])
    np.testing.assert_allclose(out["traj"].cpu().numpy()[:, :64], ref["traj"][:, :64], atol=1e-7)
    with pytest.raises(ValueError):
        engine.move_ik_plan(q0, tg, engine.ik_params(), traj_cap=1)


def test_planner_batch_vs_c_oracle_fp64(golden, oracle_chain):
    """2048 random shelf-box moves (BASELINE cfg4 distribution) through the whole state machine."""
    n = 1024
    w = synthetic.waypoint_envs(n, seed=11, dtype=torch.float64)
    cnt = torch.zeros(4, dtype=torch.int64, device="cuda")
    out = engine.move_ik_plan(w["q_start"].cuda(), w["goal"].cuda(), engine.ik_params(), counters=cnt)
    # same bound as the device default (4*200+64): part of the shelf box is out of reach and the
    # unbounded reference loop would never return there
    ref = c_oracle.move_plan(oracle_chain, w["q_start"].numpy(), w["goal"].numpy(), max_outer=864, nthreads=8)
    tl = out["traj_len"].cpu().numpy()
    same = tl == ref["traj_len"]
    assert same.mean() > 0.995
    traj = out["traj"].cpu().numpy()
    for k in np.nonzero(same)[0][:256]:
        np.testing.assert_allclose(traj[k, : tl[k]], ref["traj"][k, : tl[k]], atol=1e-8)
    assert int(cnt[0]) == int(out["n_solves"].sum())


def test_value_type_planner_pair_equals_lane_and_follows_the_reference(golden, oracle_chain):
    """move_ik_plan_v_kernel<float> (what FP32 on the specialised tree runs) and <F2> (two envs per lane,
    packed FP32): the same operations per env, so every output is bit-identical - including envs that go
    through the fallback strategies, run into the round bound, or overflow the trajectory capacity.
    Against the reference's own trajectories both keep the waypoint counts on ordinary moves and every
    waypoint within 1e-4 m; against the scalar-template kernel (generic tree) the structure agrees."""
    for n, seed in ((1, 0), (65, 1), (3000, 2), (70_001, 3)):
        w = synthetic.waypoint_envs(n, seed=seed, device="cuda")   # part of the shelf box is out of reach
        goal = w["goal"].clone()
        goal[::41] = torch.tensor([2.5, 0.0, 0.5], device="cuda")  # plus some hopeless goals
        kw = dict(max_outer=40, traj_cap=96)
        outs = {}
        for kin in ("spec_lane", "spec_pair", "generic"):
            cnt = torch.zeros(4, dtype=torch.int64, device="cuda")
            outs[kin] = (engine.move_ik_plan(w["q_start"], goal, engine.ik_params(kinematics=kin), counters=cnt, **kw), cnt)
        (a, ca), (b, cb), (g, cg) = outs["spec_lane"], outs["spec_pair"], outs["generic"]
        for f in ("traj_len", "n_solves", "status", "q_final", "traj"):
            assert torch.equal(a[f], b[f]), (n, f)
        assert torch.equal(ca, cb) and int(ca[0]) == int(a["n_solves"].sum())
        assert int((a["status"] & 2).ne(0).sum()) >= (n + 40) // 41 - 1     # the hopeless goals hit the round bound
        same = (a["traj_len"] == g["traj_len"]) & (a["status"] == g["status"])
        assert float(same.float().mean()) > 0.98
    g = golden
    ref = c_oracle.move_plan(oracle_chain, g["q_start"], g["target"], traj_cap=g["traj"].shape[1])
    clean = ref["n_solves"] < ref["traj_len"]
    for kin in ("spec_lane", "spec_pair"):
        out = engine.move_ik_plan(torch.tensor(g["q_start"], dtype=torch.float32, device="cuda"),
                                  torch.tensor(g["target"], dtype=torch.float32, device="cuda"),
                                  engine.ik_params(kinematics=kin), traj_cap=g["traj"].shape[1])
        tl, traj = out["traj_len"].cpu().numpy(), out["traj"].double().cpu().numpy()
        for k in np.nonzero(clean)[0]:
            L = int(g["traj_len"][k])
            assert tl[k] == L, (kin, k)
            assert np.abs(traj[k, :L] - g["traj"][k, :L]).max() < 1e-4
        for k in np.nonzero(~clean)[0]:
            assert tl[k] == 202 and int(out["status"][k]) == 0


def test_longest_plan_first_order_is_a_permutation_and_changes_nothing_but_the_schedule():
    """pnp_move_plan_order: env indices by descending |goal - FK(q_start)| (the length of a MoveIKSkill plan
    follows it, move.py:106-137).  The planner's outputs are indexed by env, so every one of them must be
    bit-identical with and without the order - f32 value-type kernels, the generic-tree kernel and f64."""
    for n, seed in ((1, 0), (33, 1), (1025, 2), (40_000, 3)):
        w = synthetic.waypoint_envs(n, seed=seed, device="cuda")
        goal = w["goal"].clone()
        goal[::57] = torch.tensor([2.5, 0.0, 0.5], device="cuda")
        for dtype, kins in ((torch.float32, ("spec_lane", "spec_pair", "generic")), (torch.float64, ("specialized",))):
            qs, gl = w["q_start"].to(dtype), goal.to(dtype)
            for kin in kins:
                order = engine.move_plan_order(qs, gl, kin)
                o = order.cpu().numpy().astype(np.int64)
                assert np.array_equal(np.sort(o), np.arange(n)), (n, kin)
                p0 = engine.fk_jac(qs, want_quat=False, want_jac=False, kinematics="generic" if kin == "generic" else "auto")[0]
                d0 = (gl - p0).norm(dim=1).cpu().numpy()[o]
                bucket = np.minimum((d0 * 64.0).astype(np.int64), 127)
                assert np.all(np.diff(bucket) <= 1)   # descending up to the float rounding of a bucket edge
                assert np.all(np.diff(bucket.astype(np.float64)).cumsum() <= 1)
                pk = engine.ik_params(kinematics=kin)
                kw = dict(max_outer=30, traj_cap=80)
                ca, cb = (torch.zeros(4, dtype=torch.int64, device="cuda") for _ in range(2))
                a = engine.move_ik_plan(qs, gl, pk, counters=ca, order=None, **kw)
                b = engine.move_ik_plan(qs, gl, pk, counters=cb, order=order, **kw)
                for f in ("traj_len", "n_solves", "status", "q_final", "traj"):
                    assert torch.equal(a[f], b[f]), (n, kin, f)
                assert torch.equal(ca, cb)
    # "auto" orders from 2^16 envs up (one call: PnpMoveParams.compute_order) and reuses the buffers of a
    # previous call when asked to
    n = engine.PLAN_ORDER_MIN
    w = synthetic.waypoint_envs(n, seed=9, device="cuda")
    pk = engine.ik_params()
    a = engine.move_ik_plan(w["q_start"], w["goal"], pk, order=None, traj_cap=64, max_outer=30)
    keep = {k: v.clone() for k, v in a.items()}
    b = engine.move_ik_plan(w["q_start"], w["goal"], pk, traj_cap=64, max_outer=30, out=a)
    assert b["traj"].data_ptr() == a["traj"].data_ptr()
    for f in ("traj_len", "n_solves", "status", "q_final", "traj"):
        assert torch.equal(keep[f], b[f]), f
    with pytest.raises(ValueError):
        engine.move_ik_plan(w["q_start"], w["goal"], pk, traj_cap=32, out=a)
    with pytest.raises(ValueError):
        engine.move_ik_plan(w["q_start"], w["goal"], pk, order=torch.zeros(3, dtype=torch.int32, device="cuda"))


def test_planner_fused_pass_is_bit_identical_to_separate_passes(monkeypatch):
    """The pass that accepts a solve (or takes FK(q_start)) is also the first iteration of the next solve
    (same q_current, hence the same FK and Jacobian; only the target changes, move.py:128-137): one pass
    less per solve and bit-identical outputs against PNP_WAYPOINT_FUSE=0, which spends a separate pass -
    with fallback strategies, hopeless goals, iteration caps of 1 and 2, and starts outside the joint limits."""
    for n, kw_ik in ((1, {}), (67, {}), (5000, {}), (5000, dict(max_iters=1)), (5000, dict(max_iters=2)),
                     (5000, dict(pos_thresh=5e-3, damping=0.05)), (70_000, {})):
        w = synthetic.waypoint_envs(n, seed=n % 11 + 2, device="cuda")
        goal = w["goal"].clone()
        goal[::43] = torch.tensor([2.5, 0.0, 0.5], device="cuda")
        q0 = w["q_start"].clone()
        q0[::5, 0] = 3.1
        q0[1::9, 3] = 0.2
        for kin in ("spec_lane", "spec_pair"):
            outs = []
            for fuse in ("0", "1"):
                monkeypatch.setenv("PNP_WAYPOINT_FUSE", fuse)
                cnt = torch.zeros(4, dtype=torch.int64, device="cuda")
                outs.append((engine.move_ik_plan(q0, goal, engine.ik_params(kinematics=kin, **kw_ik), counters=cnt,
                                                 max_outer=40, traj_cap=96), cnt))
            (a, ca), (b, cb) = outs
            for f in ("traj_len", "n_solves", "status", "q_final", "traj"):
                assert torch.equal(a[f], b[f]), (n, kw_ik, kin, f)
            assert torch.equal(ca, cb) and int(ca[0]) > 0
    monkeypatch.delenv("PNP_WAYPOINT_FUSE")


def test_planner_full_size_properties():
    """The bench's planner workload (2^20 reachable plans, one launch, longest plan first) through properties that
    need no oracle (skills/move.py:91-191): the trajectory starts at FK(q_start); every step between accepted
    points is at most step_size + the accept tolerance; the plan ends within pos_thresh of the goal, the goal
    itself being appended when the last accepted point is further away; q_final reproduces the last accepted
    point under FK; the solve counters agree with the per-env counts; no env breaks, caps or overflows."""
    n = 1 << 20
    tree = KinematicTree.from_mjcf()
    engine.set_tree(tree)
    w = synthetic.reachable_move_envs(n, tree.lower, tree.upper, seed=11, device="cuda")
    goal = engine.fk_jac(w["q_goal"], want_quat=False, want_jac=False)[0]
    cnt = torch.zeros(4, dtype=torch.int64, device="cuda")
    out = engine.move_ik_plan(w["q_start"], goal, engine.ik_params(), counters=cnt)   # order="auto": sorted
    L, traj, st = out["traj_len"].long(), out["traj"], out["status"]
    assert int(st.ne(0).sum()) == 0
    assert int(L.min()) >= 1 and int(L.max()) <= 202
    assert int(cnt[0]) == int(out["n_solves"].sum()) and int(cnt[1]) <= int(cnt[0])
    start = engine.fk_jac(w["q_start"], want_quat=False, want_jac=False)[0]
    assert float((traj[:, 0] - start).abs().max()) < 1e-6
    idx = torch.arange(traj.shape[1], device="cuda")[None, :]
    seg = (traj[:, 1:] - traj[:, :-1]).norm(dim=2)
    inner = idx[:, 1:] < (L[:, None] - 1)          # steps between accepted points (the last step may be the appended goal)
    assert float(seg[inner].max()) < 0.01 + 0.02 + 1e-4   # step_size + accept tolerance (move.py:131)
    last = traj[torch.arange(n, device="cuda"), L - 1]
    d_last = (last - goal).norm(dim=1)
    assert float(d_last.max()) <= 0.01 + 1e-6              # pos_thresh (move.py:106) - or exactly the goal:
    appended = (last == goal).all(dim=1)
    fk_final = engine.fk_jac(out["q_final"], want_quat=False, want_jac=False)[0]
    prev = traj[torch.arange(n, device="cuda"), (L - 2).clamp(min=0)]
    last_accepted = torch.where(appended[:, None] & (L[:, None] > 1), prev, last)
    assert float((fk_final - last_accepted).abs().max()) < 2e-5
    # the same batch in index order: bit-identical (the schedule is not part of the result)
    ref = engine.move_ik_plan(w["q_start"], goal, engine.ik_params(), order=None)
    for f in ("traj_len", "n_solves", "status", "q_final"):
        assert torch.equal(out[f], ref[f]), f
    mask = idx < L[:, None]
    assert torch.equal(traj[mask], ref["traj"][mask])


def test_fallback_strategy_2_is_exercised_against_the_oracle(oracle_chain):
    """move.py:160-176 (strategy 2: step with y frozen) cannot be reached with the reference's own constants - strategy
    1 re-solves for a point <= 1 mm away with pos_thresh = 1e-3 and always "succeeds"  - so the 18
    reference-generated moves never enter it.  With an IK that is allowed a single pass and a tighter threshold,
    strategy 1 does fail and the planner walks through strategy 2, its success (`continue`) and its failure (`break`,
    status bit 1): the FP64 device planner must reproduce the C oracle's state machine exactly, and the FP32 value-type
    kernels (one / two envs per lane) must agree with each other bit for bit and with FP64 on nearly every env."""
    n = 2048
    w = synthetic.waypoint_envs(n, seed=11, dtype=torch.float64)
    for thr, step in ((8e-4, 0.05), (3e-4, 0.01)):
        p64 = engine.ik_params(max_iters=1, pos_thresh=thr)
        kw = dict(step_size=step, max_outer=200, traj_cap=256)
        ref = c_oracle.move_plan(oracle_chain, w["q_start"].numpy(), w["goal"].numpy(), max_iters=1, ik_pos_thresh=thr,
                                 nthreads=8, **kw)
        assert (ref["status"] & 1).sum() > n // 2          # most envs end in the `break` after strategy 2 fails,
        assert (ref["status"] & 2).sum() > 0               # the rest keep going on fallback successes up to the round bound
        out = engine.move_ik_plan(w["q_start"].cuda(), w["goal"].cuda(), p64, **kw)
        np.testing.assert_array_equal(out["status"].cpu().numpy(), ref["status"])
        np.testing.assert_array_equal(out["traj_len"].cpu().numpy(), ref["traj_len"])
        np.testing.assert_array_equal(out["n_solves"].cpu().numpy(), ref["n_solves"])
        tl = ref["traj_len"]
        traj = out["traj"].cpu().numpy()
        for k in range(0, n, 7):
            np.testing.assert_allclose(traj[k, : tl[k]], ref["traj"][k, : tl[k]], atol=1e-8)
        qs, gl = w["q_start"].float().cuda(), w["goal"].float().cuda()
        a = engine.move_ik_plan(qs, gl, engine.ik_params(max_iters=1, pos_thresh=thr, kinematics="spec_lane"), **kw)
        b = engine.move_ik_plan(qs, gl, engine.ik_params(max_iters=1, pos_thresh=thr, kinematics="spec_pair"), **kw)
        for f in ("traj_len", "n_solves", "status", "q_final", "traj"):
            assert torch.equal(a[f], b[f]), f
        same = (a["status"].cpu().numpy() == ref["status"]) & (a["traj_len"].cpu().numpy() == ref["traj_len"])
        assert same.mean() > 0.9, same.mean()


def test_staged_trajectory_stores_equal_direct_stores_and_stay_inside_their_rows():
    """The FP32 planner stages the points of accepted solves in shared memory and writes them four at a time (trajectory
    capacity % 4 == 0, 160-thread blocks); any other capacity takes the kernel that stores every point directly.  Same plans,
    same points; nothing is written past traj_len or past the capacity (rows pre-filled with a sentinel), also when plans
    overflow a small capacity (status bit 4) and when goals are unreachable."""
    n = 50_000
    mv = synthetic.reachable_move_envs(n, KinematicTree.from_mjcf().lower, KinematicTree.from_mjcf().upper, seed=4, device="cuda")
    goal = engine.fk_jac(mv["q_goal"], want_quat=False, want_jac=False)[0]
    goal[::301] = torch.tensor([2.5, 0.0, 0.5], device="cuda")  # unreachable: runs into max_outer
    pk = engine.ik_params()
    runs = {}
    for cap in (256, 254, 12, 10):  # 256 / 12: staged; 254 / 10: direct
        out = dict(traj=torch.full((n, cap, 3), -7.0, device="cuda"), traj_len=torch.empty(n, dtype=torch.int32, device="cuda"),
                   q_final=torch.empty((n, 7), device="cuda"), n_solves=torch.empty(n, dtype=torch.int32, device="cuda"),
                   status=torch.empty(n, dtype=torch.int32, device="cuda"))
        for order in ("auto", None):
            r = engine.move_ik_plan(mv["q_start"], goal, pk, traj_cap=cap, max_outer=40, out=out, order=order)
            tl = r["traj_len"].long()
            stored = torch.minimum(tl, torch.tensor(cap, device="cuda"))
            idx = torch.arange(cap, device="cuda")[None, :]
            untouched = r["traj"][idx.expand(n, cap) >= stored[:, None]]
            assert bool((untouched == -7.0).all()), (cap, order)
            assert bool((r["traj"][idx.expand(n, cap) < stored[:, None]] != -7.0).all()), (cap, order)
            assert torch.equal((r["status"] & 4) != 0, tl > cap), (cap, order)
            runs[(cap, order)] = {k: v.clone() for k, v in r.items() if not k.startswith("_")}
            out["traj"].fill_(-7.0)
    ref = runs[(256, None)]
    for key, r in runs.items():
        cap = key[0]
        for f in ("traj_len", "n_solves", "q_final"):
            assert torch.equal(ref[f], r[f]), (key, f)
        assert torch.equal(ref["status"] & 3, r["status"] & 3), key
        m = min(cap, 254)
        keep = torch.arange(m, device="cuda")[None, :] < torch.minimum(ref["traj_len"].long(), torch.tensor(m, device="cuda"))[:, None]
        assert torch.equal(ref["traj"][:, :m][keep], r["traj"][:, :m][keep]), key
