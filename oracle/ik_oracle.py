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
    """Restated controller.  Attribute names follow the reference (ik_solver.py:27-33)."""

    def __init__(self, model, data, site_name: str = "ee_center_site"):
        self.model, self.data = model, data
        self.site_id = model.site(site_name).id  # :30
        self.joint_ids = np.arange(7)  # :31
        rng = model.jnt_range[:7]
        self.lower, self.upper = rng[:, 0].copy(), rng[:, 1].copy()  # :32-33

    # -- pieces of the loop body, in the reference's arithmetic order ------------------------
    def _write_q(self, q):
        self.data.qpos[:7] = q  # :51 / :82
        mujoco.mj_forward(self.model, self.data)  # :52 / :83

    def _site_pos(self):
        return self.data.site_xpos[self.site_id].copy()  # :59 / :88

    def _dls_update(self, q, residual, damping, step_limit):
        nv = self.model.nv
        jac_p, jac_r = np.zeros((3, nv)), np.zeros((3, nv))  # :70-71 (rotational part unused)
        mujoco.mj_jacSite(self.model, self.data, jac_p, jac_r, self.site_id)  # :72
        jac = jac_p[:3, :]  # :74
        gram = jac @ jac.T + damping * np.eye(3)  # :79 (damping un-squared)
        dq_all = jac.T @ np.linalg.solve(gram, residual[:3])  # :75, :78-79
        dq = np.clip(dq_all[:7], -step_limit, step_limit)  # :80
        return np.clip(q + dq, self.lower, self.upper)  # :81

    def solve(self, target_pos, q_init, max_iters=100, pos_thresh=1e-3, damping=1e-2, step_limit=0.1):
        q = np.array(q_init, dtype=np.float64).copy()  # :50
        self._write_q(q)  # :51-52
        hit, n_iter = False, 0  # :54-55
        for k in range(max_iters):  # :57
            mujoco.mj_kinematics(self.model, self.data)  # :58
            residual = target_pos - self._site_pos()  # :59-60
            if np.linalg.norm(residual) < pos_thresh:  # :61, :64 (tested BEFORE the update)
                hit, n_iter = True, k + 1  # :65-66
                break  # :67
            q = self._dls_update(q, residual, damping, step_limit)  # :70-81
            self._write_q(q)  # :82-83
            n_iter = k + 1  # :85
        end_pos = self._site_pos()  # :88
        end_err = np.linalg.norm(end_pos - target_pos)  # :89
        ok = bool(hit and end_err < pos_thresh * 2)  # :92
        return IKResult(ok, q.copy(), end_pos.copy(), float(end_err), int(n_iter), bool(hit))  # :94-101


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
