"""ORACLE (test infrastructure): FP64 restatement of MoveIKSkill.reset's trajectory planner.

Follows /root/reference/panda_mujoco_gym/skills/move.py:76-191 line by line (adaptive step,
accept rule, the double failure increment, fallback strategies 1-3, final-point append), with the
IK solves going to oracle/ik_oracle.py.  Orientation bookkeeping (quat_traj) is constant in the
reference (move.py:134) and is not reproduced.

Pinning: oracle/gen_golden.py runs the reference's *own* MoveIKSkill.reset (move.py imported
unmodified through oracle/ref_harness.py, stub env over the restated engine) and this
restatement on the same cases; tests/test_oracle.py requires bit-identical trajectories.
"""

from __future__ import annotations

import numpy as np

from . import ik_oracle, mj_oracle


def plan(model, q_start, target_pos, pos_thresh=0.01, max_traj_points=200, step_size=0.01, max_outer=None):
    """Returns dict(pos_traj [L,3], q_final [7], n_solves, broke)."""
    data = mj_oracle.MjData(model)
    data.qpos[:7] = q_start
    mj_oracle.mj_forward(model, data)
    sid = model.site("ee_center_site").id
    target_pos = np.asarray(target_pos, float)  # move.py:69
    tmp = mj_oracle.MjData(model)  # :84 deepcopy(data)
    tmp.qpos[:] = data.qpos
    ctl = ik_oracle.JacobianIKController(model, tmp)  # :85
    pos_traj = []
    start_pos = data.site_xpos[sid].copy()  # :91
    q_current = data.qpos[:7].copy()  # :93
    pos_current = start_pos.copy()
    pos_traj.append(pos_current.copy())  # :98
    point_count = 0
    consecutive_failures = 0
    max_consecutive_failures = 3
    n_solves, broke, outer = 0, False, 0
    while np.linalg.norm(pos_current - target_pos) > pos_thresh and point_count < max_traj_points:  # :106-107
        outer += 1
        if max_outer is not None and outer > max_outer:
            break
        direction = target_pos - pos_current  # :110
        distance = np.linalg.norm(direction)  # :111
        adaptive_step = min(step_size, distance * 0.1)  # :114
        max_step_size = 0.02
        adaptive_step = min(adaptive_step, max_step_size)  # :117
        if consecutive_failures > 0:  # :118-119
            adaptive_step *= 0.5
        if distance > adaptive_step:  # :122-125
            next_pos = pos_current + direction * adaptive_step / distance
        else:
            next_pos = target_pos.copy()
        ik_result = ctl.solve(next_pos, q_current)  # :128
        n_solves += 1
        if ik_result.success and ik_result.pos_error < step_size * 2:  # :131
            pos_traj.append(ik_result.final_pos.copy())
            pos_current = ik_result.final_pos.copy()
            q_current = ik_result.q.copy()
            consecutive_failures = 0
        else:
            consecutive_failures += 1  # :142
            if consecutive_failures >= max_consecutive_failures:  # :144
                smaller_step = adaptive_step * 0.1  # :149
                if distance > smaller_step:
                    fallback_pos = pos_current + direction * smaller_step / distance
                    fallback_result = ctl.solve(fallback_pos, q_current)  # :152
                    n_solves += 1
                    if fallback_result.success:
                        pos_traj.append(fallback_result.final_pos.copy())
                        pos_current = fallback_result.final_pos.copy()
                        q_current = fallback_result.q.copy()
                        consecutive_failures = 0
                        continue
                alt_direction = direction.copy()  # :163-164
                alt_direction[1] = 0
                if np.linalg.norm(alt_direction) > 0.001:
                    alt_direction = alt_direction / np.linalg.norm(alt_direction)
                    alt_pos = pos_current + alt_direction * adaptive_step
                    alt_result = ctl.solve(alt_pos, q_current)  # :168
                    n_solves += 1
                    if alt_result.success:
                        pos_traj.append(alt_result.final_pos.copy())
                        pos_current = alt_result.final_pos.copy()
                        q_current = alt_result.q.copy()
                        consecutive_failures = 0
                        continue
                broke = True  # :178-180
                break
            else:
                consecutive_failures += 1  # :183 (double increment)
                continue
        point_count += 1  # :186
    if np.linalg.norm(pos_current - target_pos) > pos_thresh:  # :189-191
        pos_traj.append(target_pos.copy())
    return dict(pos_traj=np.array(pos_traj), q_final=q_current, n_solves=n_solves, broke=broke)
