"""ORACLE (test infrastructure) for the pose-mode IK EXTENSION.

There is no reference arithmetic for this: FrankaEnv.solve_ik(target_pos, target_quat, q_init)
(/root/reference/panda_mujoco_gym/envs/panda_env.py:399-409) imports
``panda_mujoco_gym.skills.ik_solver.solve_ik``, which does not exist.  This module therefore states
the semantics the GPU kernel implements - the reference's position loop (ik_solver.py:50-101)
widened to a 6-row task - in FP64 NumPy on the restated engine (mj_oracle.py), so that the kernel
can be checked iteration for iteration.  PARITY UNPINNED (no reference behaviour exists).

    e = [target_pos - p ; w * quat2vel(target_quat (x) conj(site_quat))]   (mju_quat2Vel, dt = 1)
    J = [jacp ; w * jacr][:, :7];  dq = J^T solve(J J^T + damping I6, e)
    clip(dq, +-step_limit); q = clip(q + dq, lower, upper)
    converged when |e_pos| < pos_thresh and |rotvec| < rot_thresh, tested before the update
"""

from __future__ import annotations

import math

import numpy as np

from . import mj_oracle as mujoco


def quat2vel(q):
    axis = np.array(q[1:4], dtype=np.float64)
    sin_a_2 = math.sqrt(axis[0] ** 2 + axis[1] ** 2 + axis[2] ** 2)
    speed = 2 * math.atan2(sin_a_2, q[0])
    if speed > math.pi:
        speed -= 2 * math.pi
    return axis * (speed / sin_a_2) if sin_a_2 > 0 else np.zeros(3)


def solve_pose(model, data, target_pos, target_quat, q_init, max_iters=100, pos_thresh=1e-3, rot_thresh=1e-2,
               damping=1e-2, step_limit=0.1, rot_weight=1.0, site_name="ee_center_site"):
    sid = model.site(site_name).id
    lower, upper = model.jnt_range[:7, 0], model.jnt_range[:7, 1]
    tq = np.asarray(target_quat, dtype=np.float64)
    tq = tq / np.linalg.norm(tq)
    q = np.array(q_init, dtype=np.float64)
    converged, iterations = False, 0
    pe = re = 0.0
    for i in range(max_iters + 1):
        data.qpos[:7] = q
        mujoco.mj_forward(model, data)
        p = data.site_xpos[sid].copy()
        qc = np.empty(4)
        mujoco.mju_mat2Quat(qc, data.site_xmat[sid])
        conj = np.array([qc[0], -qc[1], -qc[2], -qc[3]])
        rv = quat2vel(mujoco.mju_mulQuat(tq, conj))
        e_pos = np.asarray(target_pos, dtype=np.float64) - p
        pe, re = float(np.linalg.norm(e_pos)), float(np.linalg.norm(rv))
        if i == max_iters:
            iterations = max_iters
            break
        if pe < pos_thresh and re < rot_thresh:
            converged, iterations = True, i + 1
            break
        jp, jr = np.zeros((3, model.nv)), np.zeros((3, model.nv))
        mujoco.mj_jacSite(model, data, jp, jr, sid)
        J = np.vstack([jp[:, :7], rot_weight * jr[:, :7]])
        e = np.concatenate([e_pos, rot_weight * rv])
        dq = J.T @ np.linalg.solve(J @ J.T + damping * np.eye(6), e)
        q = np.clip(q + np.clip(dq, -step_limit, step_limit), lower, upper)
    success = converged and pe < 2 * pos_thresh and re < 2 * rot_thresh
    return dict(success=success, converged=converged, q=q, final_pos=p, final_quat=qc, pos_error=pe, rot_error=re,
                iterations=iterations)
