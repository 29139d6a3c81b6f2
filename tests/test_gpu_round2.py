"""GPU: the round-2 additions - compact IK records, the small-batch latency kernel, per-stream launch scratch, the single-row reward mailbox, plan-order validation and the
checks on caller-supplied output buffers.  Everything goes through the C ABI (ctypes)."""
import threading

import numpy as np
import pytest
import torch

from conftest import NEUTRAL
from mujoco_panda_pnp_b200 import KinematicData, KinematicTree, engine, synthetic
from mujoco_panda_pnp_b200.envs import FrankaShelfPNPReward
from mujoco_panda_pnp_b200.skills import JacobianIKController
from oracle import c_oracle

pytestmark = pytest.mark.gpu

FIELDS = ("q", "final_pos", "pos_error", "iterations", "converged", "success")


@pytest.fixture(scope="module")
def tree(cuda_lib):
    t = KinematicTree.from_mjcf()
    engine.set_tree(t)
    return t


def _targets(tree, n, seed):
    q = synthetic.random_joint_configs(n, tree.lower, tree.upper, seed=seed, device="cuda")
    return engine.fk_jac(q, want_quat=False, want_jac=False)[0]


def _neutral():
    return torch.tensor(NEUTRAL, dtype=torch.float32, device="cuda")


@pytest.mark.parametrize("n", [1, 31, 32, 33, 4096, 148 * 128])
def test_small_batch_latency_kernel_equals_refill_kernel(tree, n):
    """AUTO sends batches of <= 128 queries per SM to ik_solve_small_kernel (no ticket, single-basic-block loop);
    spec_lane forces the persistent refill kernel.  Same arithmetic: every output and the counters are bit-identical,
    for broadcast and per-query q_init, packed / separate / compact layouts, unreachable targets and a q_init
    outside the joint limits that converges on its first pass (returned untouched, ik_solver.py:61-67)."""
    targets = _targets(tree, n, seed=11 + n)
    targets[::29] = torch.tensor([2.5, 0.0, 0.5], device="cuda")  # unreachable: runs out of iterations
    q0 = (_neutral() + 0.2 * torch.randn((n, 7), device="cuda")).contiguous()
    q0[::13, 0] = 3.2  # above joint 1's upper limit (2.8973)
    t0 = targets.clone()
    t0[::13] = engine.fk_jac(q0[::13].contiguous(), want_quat=False, want_jac=False)[0]
    for qi, tg in ((_neutral(), targets), (q0, t0)):
        for packed in (True, False):
            ca = torch.zeros(4, dtype=torch.int64, device="cuda")
            cb = torch.zeros(4, dtype=torch.int64, device="cuda")
            a = engine.ik_solve(tg, qi, engine.ik_params(kinematics="spec_lane"), packed=packed, counters=ca)
            b = engine.ik_solve(tg, qi, engine.ik_params(kinematics="auto"), packed=packed, counters=cb)
            for f in FIELDS:
                assert torch.equal(getattr(a, f), getattr(b, f)), (n, f, packed)
            assert torch.equal(ca, cb) and int(ca[0]) == n and int(ca[3]) == int(a.iterations.sum())
        c = engine.ik_solve(tg, qi, engine.ik_params(), compact=True)
        for f in ("q", "iterations", "converged", "success"):
            assert torch.equal(getattr(a, f), getattr(c, f)), (n, f, "compact")
    sel = torch.arange(0, n, 13, device="cuda")
    first = engine.ik_solve(t0, q0, engine.ik_params())
    assert bool((first.iterations[sel] == 1).all()) and torch.equal(first.q[sel], q0[sel])
    # max_iters = 0: nothing converges, zero iterations, q_init comes back
    z = engine.ik_solve(targets, _neutral(), engine.ik_params(max_iters=0))
    assert int(z.iterations.max()) == 0 and not bool(z.converged.any())
    assert torch.equal(z.q, _neutral().expand(n, 7))


def test_cfg2_latency_kernel_vs_oracle(tree, oracle_chain):
    """BASELINE cfg2 through the latency kernel: iteration counts against the FP64 oracle."""
    n = 4096
    qstar = synthetic.random_joint_configs(n, tree.lower, tree.upper, seed=0, dtype=torch.float64).numpy()
    th = c_oracle.fk_jac(oracle_chain, qstar, nthreads=8)[0].astype(np.float32).astype(np.float64)
    ref = c_oracle.ik_solve(oracle_chain, th, NEUTRAL, nthreads=8)
    res = engine.ik_solve(torch.tensor(th, dtype=torch.float32, device="cuda"), _neutral(), engine.ik_params())
    flips = int((res.iterations.cpu().numpy() != ref["iterations"]).sum())
    assert flips <= 8, flips
    assert int((res.converged.cpu().numpy() != ref["converged"]).sum()) <= 2


@pytest.mark.parametrize("kin", ["spec_pair", "spec_lane", "generic"])
def test_compact_records(tree, kin):
    """pnp_ik_solve_compact_f32: q0..q6 | iterations | flags << 24 in one 32-byte record, equal to the packed result."""
    n = 300_001 if kin != "generic" else 20_011
    targets = _targets(tree, n, seed=5)
    targets[::41] = torch.tensor([2.5, 0.0, 0.5], device="cuda")
    p = engine.ik_params(kinematics=kin)
    ca = torch.zeros(4, dtype=torch.int64, device="cuda")
    cb = torch.zeros(4, dtype=torch.int64, device="cuda")
    a = engine.ik_solve(targets, _neutral(), p, counters=ca)
    q8 = torch.full((n, 8), float("nan"), device="cuda")
    b = engine.ik_solve(targets, _neutral(), p, counters=cb, compact=True, out_q8=q8)
    assert b.final_pos is None and b.pos_error is None and b.q.data_ptr() == q8.data_ptr()
    for f in ("q", "iterations", "converged", "success"):
        assert torch.equal(getattr(a, f), getattr(b, f)), f
    assert torch.equal(ca, cb)
    # host operator: 32 bytes per query come back
    m = min(n, 70_001)
    th = targets[:m].cpu().numpy()
    h = engine.ik_solve_host(th, NEUTRAL.astype(np.float32), p, chunk_rows=9_999, compact=True)
    np.testing.assert_array_equal(h["q"], a.q[:m].cpu().numpy())
    np.testing.assert_array_equal(h["iterations"], a.iterations[:m].cpu().numpy())
    np.testing.assert_array_equal(h["converged"], a.converged[:m].cpu().numpy())
    np.testing.assert_array_equal(h["success"], a.success[:m].cpu().numpy())
    assert h["counters"][0] == m and h["counters"][1] == int(a.converged[:m].sum())


def test_pair_kernel_with_many_long_queries(tree):
    """Batches whose long-running queries (unreachable targets: 100 passes) are (a) rare, (b) one in seven, (c) all of
    them, and one that is too small to fill the grid: the two-queries-per-lane kernel equals the one-query-per-lane
    kernel bit for bit, counters equal the per-query outputs, nothing is lost or solved twice (n == counters[0], every
    record written)."""
    for n, stride in (((1 << 21) + 5, 997), ((1 << 21) + 5, 7), (700_000, 1), (40_000, 3)):
        targets = _targets(tree, n, seed=n % 1000)
        targets[::stride] = torch.tensor([2.5, 0.0, 0.5], device="cuda")
        ref = engine.ik_solve(targets, _neutral(), engine.ik_params(kinematics="spec_lane"))
        for kin in ("spec_pair",):
            cnt = torch.zeros(4, dtype=torch.int64, device="cuda")
            q8 = torch.full((n, 8), float("nan"), device="cuda")
            aux = torch.full((n, 4), float("nan"), device="cuda")
            got = engine.ik_solve(targets, _neutral(), engine.ik_params(kinematics=kin), counters=cnt, out_q8=q8, out_aux4=aux)
            assert not bool(torch.isnan(q8).any()) and not bool(torch.isnan(aux[:, :3]).any())
            for f in FIELDS:
                assert torch.equal(getattr(ref, f), getattr(got, f)), (n, stride, kin, f)
            c = cnt.cpu().numpy()
            assert c[0] == n and c[1] == int(got.converged.sum()) and c[3] == int(got.iterations.long().sum())


def test_launches_on_many_streams_do_not_share_scratch(tree):
    """Launch scratch (refill tickets) is per stream: 12 streams x 6 launches of ticket-drawing kernels in flight at
    once give the single-stream results.  (Round 1 handed tickets out round-robin per launch.)"""
    n = 150_000  # big enough for the refill kernels (ticket), small enough to overlap on the device
    targets = [_targets(tree, n, seed=900 + i) for i in range(4)]
    p_lane, p_pair = engine.ik_params(kinematics="spec_lane"), engine.ik_params(kinematics="spec_pair")
    want = [engine.ik_solve(t, _neutral(), p_lane) for t in targets]
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream() for _ in range(12)]
    got = []
    for rep in range(6):
        for si, st in enumerate(streams):
            with torch.cuda.stream(st):
                k = (si + rep) % 4
                cnt = torch.zeros(4, dtype=torch.int64, device="cuda")
                got.append((k, cnt, engine.ik_solve(targets[k], _neutral(), p_pair if (si + rep) % 2 else p_lane, counters=cnt)))
    torch.cuda.synchronize()
    for k, cnt, r in got:
        assert int(cnt[0]) == n
        assert torch.equal(r.q, want[k].q) and torch.equal(r.iterations, want[k].iterations)


def test_scalar_compute_reward_mailbox_is_bit_exact_and_thread_safe(cuda_lib, golden_reward, kin_model):
    """compute_reward with (3,) goals = one launch through the mapped mailbox (pnp_reward_one_host_f64): bit-exact on
    the reference's own rows for both reward types; four threads hammering the single-query IK and reward paths of the
    shared host context get the serial answers (the mailbox is locked per call)."""
    g = {k: golden_reward[k] for k in golden_reward.files}  # NpzFile reads lazily and is not thread-safe
    for rt, key in (("dense", "reward_dense"), ("sparse", "reward_sparse")):
        env = FrankaShelfPNPReward(rt)
        for i in range(256):
            info = dict(ee_pos=g["ee_pos"][i], ee_quat=g["ee_quat"][i], fingers_width=g["fingers_width"][i],
                        task_index=int(g["task_index"][i]))
            r = env.compute_reward(g["achieved_goal"][i], g["desired_goal"][i], info)
            assert isinstance(r, np.float32) and r.view(np.uint32) == g[key][i].view(np.uint32), (rt, i)
            assert int(env.last_counters[1]) == int(g["is_success"][i])
    ctl = JacobianIKController(kin_model, KinematicData(kin_model))
    grasp = [np.array(t) for t in [(1.415, 0, 0.73), (1.415, 0, 1.03), (1.415, 0, 0.43)]]
    want_ik = [ctl.solve(t, NEUTRAL) for t in grasp]
    errors = []

    def worker(tid):
        try:
            env = FrankaShelfPNPReward("dense")
            c = JacobianIKController(kin_model, KinematicData(kin_model))
            for rep in range(60):
                i = (tid * 61 + rep) % 512
                info = dict(ee_pos=g["ee_pos"][i], ee_quat=g["ee_quat"][i], fingers_width=g["fingers_width"][i],
                            task_index=int(g["task_index"][i]))
                r = env.compute_reward(g["achieved_goal"][i], g["desired_goal"][i], info)
                if r.view(np.uint32) != g["reward_dense"][i].view(np.uint32):
                    errors.append(("reward", tid, i))
                k = (tid + rep) % 3
                s = c.solve(grasp[k], NEUTRAL)
                if s.iterations != want_ik[k].iterations or not np.array_equal(s.q, want_ik[k].q):
                    errors.append(("ik", tid, k))
        except Exception as exc:  # noqa: BLE001
            errors.append(("exc", tid, repr(exc)))

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(4)]
    [t.start() for t in threads]
    [t.join() for t in threads]
    assert not errors, errors[:5]


def test_plan_order_is_validated(tree):
    n = 5000
    w = synthetic.reachable_move_envs(n, tree.lower, tree.upper, seed=2, device="cuda")
    goal = engine.fk_jac(w["q_goal"], want_quat=False, want_jac=False)[0]
    p = engine.ik_params()
    good = engine.move_plan_order(w["q_start"], goal)
    assert engine.move_plan_order_check(good) == 0
    assert engine.move_plan_order_check(torch.arange(n, dtype=torch.int32, device="cuda").flip(0)) == 0
    dup = good.clone()
    dup[10] = dup[11]
    assert engine.move_plan_order_check(dup) == 1
    oob = good.clone()
    oob[::100] = n + 12345
    assert engine.move_plan_order_check(oob) == 50
    neg = good.clone()
    neg[3] = -1  # 0xffffffff as unsigned
    assert engine.move_plan_order_check(neg) == 1
    for bad in (dup, oob, neg):
        with pytest.raises(ValueError, match="not a permutation"):
            engine.move_ik_plan(w["q_start"], goal, p, order=bad, traj_cap=64)
    # unchecked, an out-of-range entry is skipped: no fault, the envs named by valid entries are planned as usual
    ref = engine.move_ik_plan(w["q_start"], goal, p, order=None, traj_cap=64)
    out = engine.move_ik_plan(w["q_start"], goal, p, order=oob, traj_cap=64, validate_order=False)
    torch.cuda.synchronize()
    keep = torch.ones(n, dtype=torch.bool, device="cuda")
    keep[good[::100].long()] = False  # these envs lost their entry
    assert torch.equal(out["traj_len"][keep], ref["traj_len"][keep])
    assert torch.equal(out["q_final"][keep], ref["q_final"][keep])


def test_caller_supplied_outputs_are_checked(tree):
    n = 1000
    targets = _targets(tree, n, seed=1)
    p = engine.ik_params()
    for bad in (torch.empty((n, 7), device="cuda"), torch.empty((n - 1, 8), device="cuda"),
                torch.empty((n, 8), dtype=torch.float64, device="cuda"), torch.empty((n, 16), device="cuda")[:, ::2],
                torch.empty((n, 8))):
        with pytest.raises(ValueError):
            engine.ik_solve(targets, _neutral(), p, out_q8=bad)
    with pytest.raises(ValueError):
        engine.ik_solve(targets, _neutral(), p, out_aux4=torch.empty((n, 3), device="cuda"))
    with pytest.raises(ValueError):
        engine.ik_solve_host(targets.cpu().numpy(), NEUTRAL.astype(np.float32), p, out=dict(q8=np.empty((n, 7), np.float32)))
    with pytest.raises(ValueError):
        engine.ik_solve_host(targets.cpu().numpy(), NEUTRAL.astype(np.float32), p, out=dict(aux4=np.empty((n, 4), np.float64)))
    rows = synthetic.reward_rows(n, seed=1, device="cuda", dtype=torch.float32)
    args = [rows[k] for k in ("achieved_goal", "desired_goal", "ee_pos", "ee_quat", "fingers_width", "task_index")]
    with pytest.raises(ValueError):
        engine.reward(*args, engine.reward_params(), out=torch.empty(n - 1, device="cuda"))
    with pytest.raises(ValueError):
        engine.reward(*args, engine.reward_params(), out_success=torch.empty(n, dtype=torch.float64, device="cuda"))
    with pytest.raises(ValueError):
        engine.reward_host(*[a.cpu() for a in args], engine.reward_params(), out=np.empty(n + 1, np.float32))
    # HER: wrong-size outputs, overlapping views, out-of-range future indices
    obs = torch.randn((n, 25), device="cuda")
    nxt = torch.randn((n, 25), device="cuda")
    fut = torch.randint(-1, n, (n,), device="cuda", dtype=torch.int32)
    quat = torch.randn((n, 4), device="cuda")
    task = torch.zeros(n, dtype=torch.int32, device="cuda")
    rp = engine.reward_params()
    with pytest.raises(ValueError):
        engine.her_relabel(obs, nxt, fut, quat, task, rp, out_obs=torch.empty((n, 24), device="cuda"))
    big = torch.empty((2 * n + 8, 25), device="cuda")
    big[:n] = obs
    with pytest.raises(ValueError, match="overlap"):
        engine.her_relabel(big[:n], nxt, fut, quat, task, rp, out_obs=big[4:n + 4])
    ref = engine.her_relabel(obs, nxt, fut, quat, task, rp)
    fut_bad = fut.clone()
    keep = fut_bad < 0
    fut_bad[keep] = n + 7  # past the end: treated like "keep the stored goal"
    got = engine.her_relabel(obs, nxt, fut_bad, quat, task, rp)
    for a, b in zip(ref, got):
        assert torch.equal(a, b)


def test_set_tree_swap_waits_for_running_kernels(tree, oracle_chain):
    """Replacing the constant-memory tree while a solve with the old tree is still running on a side stream must not
    change that solve (pnp_set_tree synchronises the device before overwriting a different tree)."""
    n = 1 << 20
    targets = _targets(tree, n, seed=3)
    want = engine.ik_solve(targets, _neutral(), engine.ik_params(kinematics="generic"))
    torch.cuda.synchronize()
    other = KinematicTree(link_pos=tree.link_pos * 1.1, link_rot=tree.link_rot, ee_pos=tree.ee_pos, ee_rot=tree.ee_rot,
                          lower=tree.lower, upper=tree.upper, qref=tree.qref)
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        got = engine.ik_solve(targets, _neutral(), engine.ik_params(kinematics="generic"))
    engine.set_tree(other)  # while `got` is (most likely) still running
    engine.set_tree(tree)
    torch.cuda.synchronize()
    assert torch.equal(got.q, want.q) and torch.equal(got.iterations, want.iterations)


def test_back_to_back_launches_overlap_without_changing_a_bit(tree):
    """Consecutive ik_solve_v_kernel launches on one stream are programmatic dependents (the drain of launch k overlaps the
    ramp of launch k+1, two tickets alternating).  Whatever the buffers do - ping-pong, the same buffers again, the previous
    q_out read as q_init - every launch returns what it returns when the device is drained between launches."""
    import ctypes

    from mujoco_panda_pnp_b200 import _lib

    n = (1 << 20) + 4093
    p = engine.ik_params(kinematics="spec_pair")
    tgs = [_targets(tree, n, seed=100 + i) for i in range(5)]
    for tg in tgs:
        tg[::997] = torch.tensor([2.5, 0.0, 0.5], device="cuda")  # queries that run into max_iters: a long drain
    ref = []
    for tg in tgs:  # one at a time, the device idle in between
        q8, aux = torch.empty((n, 8), device="cuda"), torch.empty((n, 4), device="cuda")
        c = torch.zeros(4, dtype=torch.int64, device="cuda")
        engine.ik_solve(tg, _neutral(), p, counters=c, out_q8=q8, out_aux4=aux)
        torch.cuda.synchronize()
        ref.append((q8, aux, c))
    # ping-pong outputs, no synchronisation, one accumulating counter block
    bufs = [(torch.full((n, 8), float("nan"), device="cuda"), torch.full((n, 4), float("nan"), device="cuda")) for _ in range(2)]
    cnt = torch.zeros(4, dtype=torch.int64, device="cuda")
    got = []
    torch.cuda.synchronize()
    for i, tg in enumerate(tgs):
        engine.ik_solve(tg, _neutral(), p, counters=cnt, out_q8=bufs[i & 1][0], out_aux4=bufs[i & 1][1])
        if i >= 3:  # the last two land in different buffer sets: compare after the final sync
            got.append((i, bufs[i & 1]))
    torch.cuda.synchronize()
    for i, (q8, aux) in got:
        assert torch.equal(q8.view(torch.int32), ref[i][0].view(torch.int32)) and torch.equal(aux.view(torch.int32), ref[i][1].view(torch.int32)), i
    assert torch.equal(cnt, sum(r[2] for r in ref))
    # the same buffers twice in a row: the second launch waits (plain stream order) and its result stands
    q8, aux = bufs[0]
    engine.ik_solve(tgs[0], _neutral(), p, out_q8=q8, out_aux4=aux)
    engine.ik_solve(tgs[1], _neutral(), p, out_q8=q8, out_aux4=aux)
    torch.cuda.synchronize()
    assert torch.equal(q8.view(torch.int32), ref[1][0].view(torch.int32)) and torch.equal(aux.view(torch.int32), ref[1][1].view(torch.int32))
    # chained through the C ABI with nothing in between: launch 2 warm-starts from the q_out of launch 1
    lib = _lib.load()
    st = torch.cuda.current_stream().cuda_stream

    def raw(tg, qi, stride, q, fp, er, it, fl):
        _lib.check(lib.pnp_ik_solve_f32(tg.data_ptr(), qi.data_ptr(), stride, n, ctypes.byref(p), q.data_ptr(), fp.data_ptr(),
                                        er.data_ptr(), it.data_ptr(), fl.data_ptr(), None, st), "pnp_ik_solve_f32")

    def outs():
        return (torch.empty((n, 7), device="cuda"), torch.empty((n, 3), device="cuda"), torch.empty(n, device="cuda"),
                torch.empty(n, dtype=torch.int32, device="cuda"), torch.empty(n, dtype=torch.uint8, device="cuda"))

    neutral = _neutral()
    a1, a2, b1, b2 = outs(), outs(), outs(), outs()
    raw(tgs[2], neutral, 0, *a1)
    torch.cuda.synchronize()
    raw(tgs[3], a1[0], 7, *a2)
    torch.cuda.synchronize()
    raw(tgs[2], neutral, 0, *b1)
    raw(tgs[3], b1[0], 7, *b2)   # reads what the launch before it is still writing: must wait for all of it
    torch.cuda.synchronize()
    for x, y in zip(a2, b2):
        assert torch.equal(x.view(torch.uint8), y.view(torch.uint8))


def test_drain_handover_in_mixed_launch_sequences(tree):
    """The pair kernel hands its last running slots to a resume launch (one query per lane) and leaves; the counters of
    that hand-over alternate per stream and are zeroed by the resume launch before.  Sequences that break the alternation -
    a one-query-per-lane launch (no resume launch behind it) between two pair launches, small batches, per-query q_init,
    the three output layouts - return what each launch returns alone; nothing is lost or solved twice."""
    n = (1 << 20) + 77
    tg = [_targets(tree, n, seed=300 + i) for i in range(3)]
    for t in tg:
        t[::211] = torch.tensor([2.5, 0.0, 0.5], device="cuda")
    qi = (_neutral() + 0.1 * torch.randn((n, 7), device="cuda")).contiguous()
    seq = [("spec_pair", tg[0], _neutral(), {}), ("spec_lane", tg[1], _neutral(), {}), ("spec_pair", tg[2], qi, {}),
           ("spec_pair", tg[1], _neutral(), {"compact": True}), ("auto", tg[0][:4096].contiguous(), _neutral(), {}),
           ("spec_pair", tg[2], _neutral(), {"packed": False}), ("spec_pair", tg[0], qi, {})]
    fields = ("q", "iterations", "converged", "success")
    ref = []
    for kin, t, q0, kw in seq:  # one at a time, on the one-query-per-lane kernel, the device idle in between
        r = engine.ik_solve(t, q0, engine.ik_params(kinematics="spec_lane" if kin != "auto" else "auto"), **kw)
        torch.cuda.synchronize()
        ref.append({f: getattr(r, f).clone() for f in fields})
    for rep in range(2):
        cnts = [torch.zeros(4, dtype=torch.int64, device="cuda") for _ in seq]
        got = [engine.ik_solve(t, q0, engine.ik_params(kinematics=kin), counters=c, **kw) for (kin, t, q0, kw), c in zip(seq, cnts)]
        torch.cuda.synchronize()
        for i, (r, want, c) in enumerate(zip(got, ref, cnts)):
            for f in fields:
                assert torch.equal(getattr(r, f), want[f]), (rep, i, f)
            assert int(c[0]) == len(want["q"]) and int(c[1]) == int(want["converged"].sum()) and int(c[3]) == int(want["iterations"].long().sum()), (rep, i)


def test_random_launch_sequences_on_two_streams(tree):
    """Random sequences of IK solves of different sizes (the latency kernel, one / two queries per lane, with and without the
    drain hand-over), q_init modes and kernels on two streams, nothing synchronised in between, now and then into the previous
    launch's buffers: every launch returns what the same solve returns alone (tools/dev/dev_pdl_stress.py runs longer ones)."""
    import random

    dev = torch.device("cuda")
    rng = random.Random(3)
    sizes = [4096, 30_000, 60_000, 300_000, (1 << 20) + 3, 3_000_000]
    inputs, ref = {}, {}
    for n in sizes:
        t = _targets(tree, n, seed=n % 977)
        t[::173] = torch.tensor([2.5, 0.0, 0.5], device=dev)
        inputs[n] = (t, (_neutral() + 0.1 * torch.randn((n, 7), device=dev)).contiguous())
        for pq in (False, True):
            r = engine.ik_solve(t, inputs[n][1] if pq else _neutral(), engine.ik_params(kinematics="spec_lane" if n > 20000 else "auto"))
            torch.cuda.synchronize()
            ref[(n, pq)] = (r.q.clone(), r.iterations.clone(), r.final_pos.clone())
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    for _ in range(2):
        got = []
        for _ in range(16):
            n, pq, si = rng.choice(sizes), rng.random() < 0.3, rng.randrange(2)
            kin = rng.choice(["auto", "auto", "spec_pair", "spec_lane"]) if n > 40000 else "auto"
            t, qi = inputs[n]
            with torch.cuda.stream(streams[si]):
                if rng.random() < 0.2 and got and got[-1] is not None and got[-1][0] == n and got[-1][4] == si:
                    bufs = got[-1][3]          # into the previous launch's buffers: the library must serialise the two
                    got[-1] = None
                else:
                    bufs = (torch.empty((n, 8), device=dev), torch.empty((n, 4), device=dev))
                r = engine.ik_solve(t, qi if pq else _neutral(), engine.ik_params(kinematics=kin), out_q8=bufs[0], out_aux4=bufs[1])
                got.append((n, pq, r, bufs, si))
        torch.cuda.synchronize()
        for g in got:
            if g is not None:
                n, pq, r, _, _ = g
                q, it, fp = ref[(n, pq)]
                assert torch.equal(r.q, q) and torch.equal(r.iterations, it) and torch.equal(r.final_pos, fp), (n, pq)
