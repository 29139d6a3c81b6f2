"""ORACLE (test infrastructure): run the reference's OWN Python for the hot path.

Only usable where /root/reference exists (this container; never on the GPU box).  The
reference modules are imported *unmodified from where they lie*:

    /root/reference/panda_mujoco_gym/skills/ik_solver.py   (JacobianIKController, IKResult)
    /root/reference/panda_mujoco_gym/envs/panda_env.py     (FrankaEnv.compute_reward & co.)

Their third-party imports cannot be satisfied in this image (mujoco, gymnasium,
gymnasium_robotics are not installable), so before importing we register stand-ins
in ``sys.modules``:

    mujoco                      -> oracle/mj_oracle.py (restated engine routines)
    gymnasium.core              -> ObsType placeholder
    gymnasium_robotics.envs.robot_env.MujocoRobotEnv -> empty base class
    gymnasium_robotics.utils.rotations               -> euler2quat restated (reward_oracle.py)

and the ``panda_mujoco_gym`` / ``.skills`` / ``.envs`` packages are registered as bare
namespace modules so that their ``__init__`` files (which register gym envs) do not run.
What executes is therefore the reference's real control flow and arithmetic on top of the
restated engine: this pins the *restated control flow* (ik_oracle.py, reward_oracle.py)
against the real code, not the engine routines against real MuJoCo.
"""

from __future__ import annotations

import importlib
import os
import sys
import types

import numpy as np

from . import mj_oracle, reward_oracle

REFERENCE_ROOT = "/root/reference"


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "panda_mujoco_gym", "skills", "ik_solver.py"))


def _install_stubs() -> None:
    if "panda_mujoco_gym" in sys.modules and getattr(sys.modules["panda_mujoco_gym"], "_pnp_stub", False):
        return
    # --- mujoco ------------------------------------------------------------------------
    mj = types.ModuleType("mujoco")
    for name in ("MjModel", "MjData", "mj_forward", "mj_kinematics", "mj_jacSite", "mju_mat2Quat"):
        setattr(mj, name, getattr(mj_oracle, name))
    mj.mjtEq = types.SimpleNamespace(mjEQ_WELD=1)
    sys.modules["mujoco"] = mj
    # --- gymnasium / gymnasium_robotics -------------------------------------------------
    gym = types.ModuleType("gymnasium")
    gym_core = types.ModuleType("gymnasium.core")
    gym_core.ObsType = object
    gym.core = gym_core
    sys.modules.setdefault("gymnasium", gym)
    sys.modules.setdefault("gymnasium.core", gym_core)
    gr = types.ModuleType("gymnasium_robotics")
    gr_envs = types.ModuleType("gymnasium_robotics.envs")
    gr_robot = types.ModuleType("gymnasium_robotics.envs.robot_env")

    class MujocoRobotEnv:  # the hot-path methods never call into the base class
        pass

    gr_robot.MujocoRobotEnv = MujocoRobotEnv
    gr_utils = types.ModuleType("gymnasium_robotics.utils")
    gr_rot = types.ModuleType("gymnasium_robotics.utils.rotations")
    gr_rot.euler2quat = reward_oracle.euler2quat
    from . import obs_oracle

    gr_rot.mat2euler = obs_oracle.mat2euler
    gr_utils.rotations = gr_rot
    for name, mod in [
        ("gymnasium_robotics", gr),
        ("gymnasium_robotics.envs", gr_envs),
        ("gymnasium_robotics.envs.robot_env", gr_robot),
        ("gymnasium_robotics.utils", gr_utils),
        ("gymnasium_robotics.utils.rotations", gr_rot),
    ]:
        sys.modules.setdefault(name, mod)
    # --- bare namespace packages so the reference's __init__ files do not execute --------
    base = os.path.join(REFERENCE_ROOT, "panda_mujoco_gym")
    for pkg, sub in [("panda_mujoco_gym", ""), ("panda_mujoco_gym.skills", "skills"), ("panda_mujoco_gym.envs", "envs")]:
        mod = types.ModuleType(pkg)
        mod.__path__ = [os.path.join(base, sub)]
        mod._pnp_stub = True
        sys.modules[pkg] = mod


def reference_ik_module():
    """The reference's ik_solver module, executed from /root/reference."""
    _install_stubs()
    return importlib.import_module("panda_mujoco_gym.skills.ik_solver")


def reference_env_module():
    """The reference's panda_env module, executed from /root/reference."""
    _install_stubs()
    return importlib.import_module("panda_mujoco_gym.envs.panda_env")


def reference_xml_path() -> str:
    return os.path.join(REFERENCE_ROOT, "panda_mujoco_gym", "assets", "shelf_pnp.xml")


class RewardProbe:
    """Minimal ``self`` for calling the reference's unbound FrankaEnv reward methods.

    Supplies exactly the attributes compute_reward reads (panda_env.py:211-244): the three
    getters (hidden simulator state) and the env scalars configured by
    FrankaShelfPNPEnv (shelf_pnp.py:17-26).
    """

    def __init__(self, env_mod, reward_type="dense", n_tasks=3, initial_object_height=0.001,
                 distance_threshold=0.05, high_pick_z=0.35):
        self._cls = env_mod.FrankaEnv
        self.HORIZONTAL_QUAT = self._cls.HORIZONTAL_QUAT
        self.VERTICAL_QUAT = self._cls.VERTICAL_QUAT
        self.reward_type = reward_type
        self.task_sequence = [f"cube{i + 1}" for i in range(n_tasks)]
        self.initial_object_height = initial_object_height
        self.distance_threshold = distance_threshold
        self.high_pick_z = high_pick_z
        self.current_task_index = 0
        self._ee_pos = np.zeros(3)
        self._ee_quat = np.array([1.0, 0, 0, 0])
        self._width = 0.0

    def goal_distance(self, a, b):
        return self._cls.goal_distance(self, a, b)

    def get_ee_position(self):
        return self._ee_pos

    def get_ee_orientation(self):
        return self._ee_quat

    def get_fingers_width(self):
        return self._width

    def reward(self, ag, dg, ee_pos, ee_quat, width, task_index):
        self._ee_pos = np.asarray(ee_pos, dtype=np.float64)
        self._ee_quat = np.asarray(ee_quat, dtype=np.float64)
        self._width = np.float64(width)
        self.current_task_index = int(task_index)
        return self._cls.compute_reward(self, ag, dg, {})

    def success(self, ag, dg):
        return self._cls._is_success(self, ag, dg)


class ObsProbe:
    """Minimal ``self`` for the reference's unbound ``FrankaEnv._get_obs`` (panda_env.py:279-301):
    model/data (restated engine), ``_utils`` (restated mujoco_utils), dt, goal and the current
    target object - exactly the attributes that method reads."""

    def __init__(self, env_mod, model, data, dt=0.05):
        from . import obs_oracle

        self._cls = env_mod.FrankaEnv
        self.model, self.data = model, data
        self._utils = obs_oracle.MujocoUtils
        self.dt = dt
        self.block_gripper = False
        self.current_target_object = "cube1"
        self.goal = None

    def get_fingers_width(self):
        return self._cls.get_fingers_width(self)

    def get_obs(self, current_obj, goal):
        self.current_target_object = current_obj
        self.goal = goal
        return self._cls._get_obs(self)


def reference_move_module():
    """The reference's skills/move.py (MoveIKSkill), executed from /root/reference."""
    _install_stubs()
    return importlib.import_module("panda_mujoco_gym.skills.move")


class MoveEnvStub:
    """The slice of the gym env that MoveIKSkill.reset touches (move.py:81-93): ``unwrapped``
    with model/data, get_ee_position, get_ee_orientation."""

    def __init__(self, model, q_start):
        self.model = model
        self.data = mj_oracle.MjData(model)
        self.data.qpos[:7] = q_start
        mj_oracle.mj_forward(model, self.data)
        self.unwrapped = self
        self._sid = model.site("ee_center_site").id

    def get_ee_position(self):
        return self.data.site_xpos[self._sid]

    def get_ee_orientation(self):
        quat = np.empty(4)
        mj_oracle.mju_mat2Quat(quat, self.data.site_xmat[self._sid])
        return quat


def reference_move_plan(model, q_start, target_pos, **kw):
    """Run the reference's own MoveIKSkill.reset and return its pos_traj as an array."""
    import contextlib
    import io

    move = reference_move_module()
    skill = move.MoveIKSkill(MoveEnvStub(model, q_start), np.asarray(target_pos, float), **kw)
    with contextlib.redirect_stdout(io.StringIO()):
        skill.reset()
    return np.array(skill.pos_traj)
