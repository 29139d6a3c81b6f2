"""ORACLE (test infrastructure): FP64 restatement of JacobianIKController.solve.

Follows /root/reference/panda_mujoco_gym/skills/ik_solver.py:26-101 line by line; the MuJoCo
calls go to the restated engine routines in oracle/mj_oracle.py.  Control-flow quirks kept on
purpose: convergence tested before the update, ``iterations = i + 1``,
damping added un-squared, per-joint step clip then limit clip, nv-wide ``dq`` truncated to 7.

Pinning: oracle/gen_golden.py runs the reference's *own* ik_solver.py (imported unmodified
from /root/reference with ``mujoco`` replaced by oracle/mj_oracle.py) and this restatement on
the same inputs; tests/golden/ik_reference_golden.json holds the reference-side outputs and
tests/test_oracle.py requires bit-identical agreement.  The engine routines underneath are
pinned only by ``home_wpt`` (see mj_oracle.py) -> "parity unpinned against real MuJoCo".
"""

from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import mj_oracle as mujoco


@dataclass
class IKResult:  # ik_solver.py:16-24
    success: bool
    q: np.ndarray
    final_pos: np.ndarray
    pos_error: float
    iterations: int
    converged: bool


class JacobianIKController:
    def __init__(self, model, data, site_name: str = "ee_center_site"):  # ik_solver.py:27-33
        self.model = model
        self.data = data
        self.site_id = model.site(site_name).id
        self.joint_ids = np.arange(7)
        self.lower = model.jnt_range[:7, 0].copy()
        self.upper = model.jnt_range[:7, 1].copy()

    def solve(self, target_pos, q_init, max_iters=100, pos_thresh=1e-3, damping=1e-2, step_limit=0.1):
        m, d, sid = self.model, self.data, self.site_id
        q = np.array(q_init, dtype=np.float64).copy()  # :50
        d.qpos[:7] = q  # :51
        mujoco.mj_forward(m, d)  # :52
        converged, iterations = False, 0  # :54-55
        for i in range(max_iters):  # :57
            mujoco.mj_kinematics(m, d)  # :58
            curr_pos = d.site_xpos[sid].copy()  # :59
            pos_err = target_pos - curr_pos  # :60
            pos_error_norm = np.linalg.norm(pos_err)  # :61
            if pos_error_norm < pos_thresh:  # :64
                converged = True
                iterations = i + 1
                break
            J_pos = np.zeros((3, m.nv))  # :70
            J_rot = np.zeros((3, m.nv))  # :71
            mujoco.mj_jacSite(m, d, J_pos, J_rot, sid)  # :72
            J = J_pos[:3, :]  # :74
            err = pos_err[:3]  # :75
            JT = J.T  # :78
            delta_q_full = JT @ np.linalg.solve(J @ JT + damping * np.eye(3), err)  # :79
            delta_q = np.clip(delta_q_full[:7], -step_limit, step_limit)  # :80
            q = np.clip(q + delta_q, self.lower, self.upper)  # :81
            d.qpos[:7] = q  # :82
            mujoco.mj_forward(m, d)  # :83
            iterations = i + 1  # :85
        final_pos = d.site_xpos[sid].copy()  # :88
        final_error = np.linalg.norm(final_pos - target_pos)  # :89
        success = bool(converged and final_error < pos_thresh * 2)  # :92
        return IKResult(success, q.copy(), final_pos.copy(), float(final_error), int(iterations), bool(converged))


def fk_site(model, data, q, site_name="ee_center_site"):
    """EE-site world position, 3x3 matrix, wxyz quaternion and 6 x 7 Jacobian at arm angles q."""
    sid = model.site(site_name).id
    data.qpos[:7] = np.asarray(q, dtype=np.float64)
    mujoco.mj_forward(model, data)
    jp, jr = np.zeros((3, model.nv)), np.zeros((3, model.nv))
    mujoco.mj_jacSite(model, data, jp, jr, sid)
    quat = np.empty(4)
    mujoco.mju_mat2Quat(quat, data.site_xmat[sid])
    return (
        data.site_xpos[sid].copy(),
        data.site_xmat[sid].reshape(3, 3).copy(),
        quat,
        np.vstack([jp[:, :7], jr[:, :7]]),
    )
