"""mujoco_panda_pnp_b200 - B200 (sm_100a) implementation of the mujoco-panda-pnp hot path.

Batched Franka Panda position IK (JacobianIKController.solve) and goal-conditioned
compute_reward / _is_success, as hand-written CUDA behind a C ABI (include/pnp_b200.h).
Importing this package does not load the CUDA library; the first call does and raises if
libpnp_b200.so is missing (no CPU fallback).
"""

from .mjcf import KinematicData, KinematicModel  # noqa: F401
from .tree import DEFAULT_ASSET, KinematicTree  # noqa: F401

__version__ = "0.1.0"


def __getattr__(name):
    # torch-dependent parts are imported lazily so `import mujoco_panda_pnp_b200` stays cheap
    if name in ("JacobianIKController", "IKSolver", "IKResult", "BatchIKResult"):
        from . import skills

        return getattr(skills, name)
    if name in ("FrankaRewardModel", "FrankaShelfPNPReward"):
        from . import envs

        return getattr(envs, name)
    if name == "engine":
        import importlib

        return importlib.import_module(".engine", __name__)
    raise AttributeError(name)
