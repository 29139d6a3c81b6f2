"""ORACLE (test infrastructure): FP64 restatement of MoveIKSkill.reset's trajectory planner.

Semantics of /root/reference/panda_mujoco_gym/skills/move.py:76-191 (adaptive step, accept rule,
the double failure increment, fallback strategies 1-3, final-point append), written as the explicit
state machine {NORMAL, FB1, FB2} that the device kernel uses instead of the reference's nested
if/continue/break, so that this file checks the *formulation* against the reference's own
trajectories.  IK solves go to oracle/ik_oracle.py.  Orientation bookkeeping (quat_traj) is constant
in the reference (move.py:134) and is not reproduced.

Pinning: oracle/gen_golden.py runs the reference's *own* MoveIKSkill.reset (move.py imported
unmodified through oracle/ref_harness.py, stub env over the restated engine);
tests/test_oracle.py requires this restatement to give bit-identical trajectories.

Reference quirk kept out of the golden set: for an unreachable target the loop never ends
(fallback 1 keeps succeeding with ever smaller steps without advancing point_count); ``max_outer``
bounds the NORMAL rounds here.
"""

from __future__ import annotations

import numpy as np

from . import ik_oracle, mj_oracle

NORMAL, FB1, FB2 = 0, 1, 2


def plan(model, q_start, target_pos, pos_thresh=0.01, max_traj_points=200, step_size=0.01, max_outer=None):
    """Returns dict(pos_traj [L,3], q_final [7], n_solves, broke)."""
    sid = model.site("ee_center_site").id
    goal = np.asarray(target_pos, float)  # move.py:69
    live = mj_oracle.MjData(model)
    live.qpos[:7] = q_start
    mj_oracle.mj_forward(model, live)
    scratch = mj_oracle.MjData(model)  # :84 the solver works on a private copy
    scratch.qpos[:] = live.qpos
    solver = ik_oracle.JacobianIKController(model, scratch)  # :85
    pos = live.site_xpos[sid].copy()  # :91, :94
    q = live.qpos[:7].copy()  # :93
    traj = [pos.copy()]  # :98
    accepted_points = fails = rounds = n_solves = 0
    state, broke = NORMAL, False
    step = 0.0
    while True:
        delta = goal - pos  # :110
        dist = np.linalg.norm(delta)  # :111 (same value as the loop test at :106)
        if state == NORMAL:
            if not (dist > pos_thresh and accepted_points < max_traj_points):  # :106-107
                break
            rounds += 1
            if max_outer is not None and rounds > max_outer:
                break
            step = min(min(step_size, dist * 0.1), 0.02)  # :114-117
            if fails > 0:
                step *= 0.5  # :118-119
            aim = pos + delta * step / dist if dist > step else goal.copy()  # :122-125
        elif state == FB1:
            aim = pos + delta * (step * 0.1) / dist  # :149-151
        else:
            flat = delta.copy()  # :163-167
            flat[1] = 0
            flat = flat / np.linalg.norm(flat)
            aim = pos + flat * step
        res = solver.solve(aim, q)  # :128 / :152 / :168
        n_solves += 1
        good = res.success and (res.pos_error < step_size * 2 if state == NORMAL else True)  # :131 / :154 / :170
        if good:
            traj.append(res.final_pos.copy())
            pos, q, fails = res.final_pos.copy(), res.q.copy(), 0
            if state == NORMAL:
                accepted_points += 1  # :186 (the fallback branches `continue` past it)
            state = NORMAL
            continue
        go_fb2 = False
        if state == NORMAL:
            fails += 1  # :142
            if fails >= 3:  # :144
                if dist > step * 0.1:  # :150
                    state = FB1
                else:
                    go_fb2 = True
            else:
                fails += 1  # :183 (second increment)
        elif state == FB1:
            go_fb2 = True
        else:
            broke = True  # :178-180
            break
        if go_fb2:
            flat = delta.copy()
            flat[1] = 0
            if np.linalg.norm(flat) > 0.001:  # :165
                state = FB2
            else:
                broke = True
                break
    if np.linalg.norm(pos - goal) > pos_thresh:  # :189-191
        traj.append(goal.copy())
    return dict(pos_traj=np.array(traj), q_final=q, n_solves=n_solves, broke=broke)
