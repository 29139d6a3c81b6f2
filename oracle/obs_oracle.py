"""ORACLE (test infrastructure): FP64 restatement of FrankaEnv._get_obs.

Follows /root/reference/panda_mujoco_gym/envs/panda_env.py:279-301 (+ :348-352 get_fingers_width)
line by line.  The helpers it calls live in gymnasium-robotics==1.2.2
(/root/reference/requirements.txt:3; third-party, absent from /root/reference, not installable
here), restated from the published source:

    mujoco_utils.get_site_xpos / get_site_xmat : data.site_xpos[id], data.site_xmat[id].reshape(3,3)
    mujoco_utils.get_site_xvelp / get_site_xvelr: mj_jacSite -> jacp @ qvel, jacr @ qvel
    mujoco_utils.get_joint_qpos                : qpos slice of the named joint
    rotations.mat2euler                        : the openai/mujoco-py convention (below)

``MujocoRobotEnv.dt`` = model.opt.timestep * n_substeps = 0.002 * 25 (shelf_pnp.py:19).

Pinning: oracle/gen_golden.py executes the reference's *own* ``FrankaEnv._get_obs`` (imported
unmodified, with ``self._utils`` / ``rotations`` bound to these restated helpers) and stores
its outputs in tests/golden/obs_reference_golden.npz; tests/test_oracle.py requires this module to
be bit-identical.  The helpers themselves are unpinned against the real packages.
"""

from __future__ import annotations

import numpy as np

from . import mj_oracle as mujoco

_FLOAT_EPS = np.finfo(np.float64).eps
_EPS4 = _FLOAT_EPS * 4.0


def mat2euler(mat):
    """gymnasium_robotics.utils.rotations.mat2euler."""
    mat = np.asarray(mat, dtype=np.float64)
    assert mat.shape[-2:] == (3, 3), f"Invalid shape matrix {mat}"
    cy = np.sqrt(mat[..., 2, 2] * mat[..., 2, 2] + mat[..., 1, 2] * mat[..., 1, 2])
    condition = cy > _EPS4
    euler = np.empty(mat.shape[:-1], dtype=np.float64)
    euler[..., 2] = np.where(
        condition, -np.arctan2(mat[..., 0, 1], mat[..., 0, 0]), -np.arctan2(-mat[..., 1, 0], mat[..., 1, 1])
    )
    euler[..., 1] = np.where(condition, -np.arctan2(-mat[..., 0, 2], cy), -np.arctan2(-mat[..., 0, 2], cy))
    euler[..., 0] = np.where(condition, -np.arctan2(mat[..., 1, 2], mat[..., 2, 2]), 0.0)
    return euler


class MujocoUtils:
    """The slice of gymnasium_robotics.utils.mujoco_utils that _get_obs touches."""

    @staticmethod
    def get_site_xpos(model, data, name):
        return data.site_xpos[model.site(name).id]

    @staticmethod
    def get_site_xmat(model, data, name):
        return data.site_xmat[model.site(name).id].reshape(3, 3)

    @staticmethod
    def _jac(model, data, name):
        jacp, jacr = np.zeros((3, model.nv)), np.zeros((3, model.nv))
        mujoco.mj_jacSite(model, data, jacp, jacr, model.site(name).id)
        return jacp, jacr

    @classmethod
    def get_site_xvelp(cls, model, data, name):
        return cls._jac(model, data, name)[0] @ data.qvel

    @classmethod
    def get_site_xvelr(cls, model, data, name):
        return cls._jac(model, data, name)[1] @ data.qvel

    @staticmethod
    def get_joint_qpos(model, data, name):
        j = model.joint(name).id
        adr, jt = int(model.jnt_qposadr[j]), int(model.jnt_type[j])
        ndim = {0: 7, 1: 4, 2: 1, 3: 1}[jt]
        return data.qpos[adr : adr + ndim].copy()


def set_state(model, data, q_arm, qvel_arm, fingers, obj_name, obj_pos, obj_quat, obj_vel):
    """Write a kinematic state into MjData and run the position stage (mj_forward)."""
    data.qpos[:] = model.qpos0
    data.qvel[:] = 0.0
    data.qpos[:7] = q_arm
    data.qvel[:7] = qvel_arm
    data.qpos[7:9] = fingers
    j = model.joint(f"{obj_name}_joint").id
    qa, da = int(model.jnt_qposadr[j]), int(model.jnt_dofadr[j])
    data.qpos[qa : qa + 3] = obj_pos
    data.qpos[qa + 3 : qa + 7] = obj_quat
    data.qvel[da : da + 6] = obj_vel
    mujoco.mj_forward(model, data)


def get_obs(model, data, current_obj, goal, dt=0.05, block_gripper=False):
    """panda_env.py:279-301 on an MjData that already holds the state."""
    U = MujocoUtils
    ee_pos = U.get_site_xpos(model, data, "ee_center_site").copy()  # :285
    ee_vel = U.get_site_xvelp(model, data, "ee_center_site").copy() * dt  # :286
    site = f"{current_obj}_site"  # :289
    obj_pos = U.get_site_xpos(model, data, site).copy()  # :290
    obj_rot = mat2euler(U.get_site_xmat(model, data, site)).copy()  # :291
    obj_velp = U.get_site_xvelp(model, data, site).copy() * dt  # :292
    obj_velr = U.get_site_xvelr(model, data, site).copy() * dt  # :293
    if not block_gripper:  # :295-299
        f1 = U.get_joint_qpos(model, data, "finger_joint1")  # :350
        f2 = U.get_joint_qpos(model, data, "finger_joint2")  # :351
        fingers_width = (f1 + f2).copy()  # :352, :296
        obs = np.concatenate([ee_pos, ee_vel, fingers_width, obj_pos, obj_rot, obj_velp, obj_velr])
    else:
        obs = np.concatenate([ee_pos, ee_vel, obj_pos, obj_rot, obj_velp, obj_velr])
    return {"observation": obs, "achieved_goal": obj_pos.copy(), "desired_goal": goal}  # :301
