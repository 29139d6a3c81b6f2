"""The device entry points never allocate, synchronise or touch the host (include/pnp_b200.h, conventions):
a sequence of them is capturable into a CUDA graph and replays on new buffer contents with the results of
the eager calls, bit for bit.  (A launch-bound caller - e.g. one BT tick of a few thousand envs: solve,
observe, reward - replays the whole tick with one graph launch.)"""
import numpy as np
import pytest
import torch

from mujoco_panda_pnp_b200 import KinematicTree, engine, synthetic

pytestmark = pytest.mark.gpu


def _inputs(n, seed, tree, dev):
    qstar = synthetic.random_joint_configs(n, tree.lower, tree.upper, seed=seed, device=dev)
    targets = engine.fk_jac(qstar, want_quat=False, want_jac=False)[0]
    rows = synthetic.reward_rows(n, seed=seed, device=dev, dtype=torch.float32)
    w = synthetic.waypoint_envs(n, seed=seed, device=dev)
    return targets, rows, w


def test_ik_reward_planner_sequence_replays_from_a_cuda_graph(cuda_lib):
    dev = torch.device("cuda")
    tree = KinematicTree.from_mjcf()
    engine.set_tree(tree)
    n = engine.PLAN_ORDER_MIN  # large enough for the longest-plan-first sort to be part of the captured sequence
    pk = engine.ik_params()
    rp = engine.reward_params("dense")
    neutral = torch.tensor(synthetic.NEUTRAL_Q, dtype=torch.float32, device=dev)
    keys = ("achieved_goal", "desired_goal", "ee_pos", "ee_quat", "fingers_width", "task_index")
    targets, rows, w = _inputs(n, 1, tree, dev)
    # static buffers the graph reads and writes
    s_targets, s_rows = targets.clone(), {k: rows[k].clone() for k in keys}
    s_qs, s_goal = w["q_start"].clone(), w["goal"].clone()
    q8 = torch.empty((n, 8), dtype=torch.float32, device=dev)
    aux = torch.empty((n, 4), dtype=torch.float32, device=dev)
    cnt = torch.zeros(4, dtype=torch.int64, device=dev)
    rew = torch.empty(n, dtype=torch.float32, device=dev)
    suc = torch.empty(n, dtype=torch.float32, device=dev)
    kw = dict(traj_cap=48, max_outer=24)
    plan = engine.move_ik_plan(s_qs, s_goal, pk, **kw)          # eager warm-up of every kernel in the sequence
    engine.ik_solve(s_targets, neutral, pk, counters=cnt, out_q8=q8, out_aux4=aux)
    engine.reward(*(s_rows[k] for k in keys), rp, out=rew, out_success=suc)
    torch.cuda.synchronize()

    def sequence():
        cnt.zero_()
        engine.ik_solve(s_targets, neutral, pk, counters=cnt, out_q8=q8, out_aux4=aux)
        engine.reward(*(s_rows[k] for k in keys), rp, out=rew, out_success=suc)
        engine.move_ik_plan(s_qs, s_goal, pk, out=plan, **kw)

    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        sequence()
    for seed in (2, 3):
        targets, rows, w = _inputs(n, seed, tree, dev)
        s_targets.copy_(targets); s_qs.copy_(w["q_start"]); s_goal.copy_(w["goal"])
        for k in keys:
            s_rows[k].copy_(rows[k])
        g.replay()
        torch.cuda.synchronize()
        got = [t.clone() for t in (q8, aux, cnt, rew, suc, plan["traj_len"], plan["q_final"], plan["n_solves"], plan["status"])]
        got_traj = plan["traj"].clone()
        sequence()                                               # the same calls, eagerly, on the same contents
        torch.cuda.synchronize()
        want = (q8, aux, cnt, rew, suc, plan["traj_len"], plan["q_final"], plan["n_solves"], plan["status"])
        for a, b in zip(got, want):
            assert torch.equal(a, b)
        L = plan["traj_len"].clamp(max=kw["traj_cap"])
        mask = torch.arange(kw["traj_cap"], device=dev)[None, :] < L[:, None]
        assert torch.equal(got_traj[mask], plan["traj"][mask])
        assert int(cnt[0]) == n and int(cnt[1]) > 0.99 * n       # and the solves did run: ~99.8 % converge
