"""Env configuration scalars of the shelf task
(/root/reference/panda_mujoco_gym/envs/shelf_pnp.py:17-26)."""

from .panda_env import FrankaRewardModel

SHELF_PNP_CONFIG = dict(
    n_substeps=25,
    block_gripper=False,
    distance_threshold=0.05,
    obj_x_range=0.02,
    obj_y_range=0.2,
)


class FrankaShelfPNPReward(FrankaRewardModel):
    """Reward arithmetic of FrankaShelfPNPEnv(reward_type) (shelf_pnp.py:11-26)."""

    def __init__(self, reward_type, **kwargs):
        kwargs.setdefault("distance_threshold", SHELF_PNP_CONFIG["distance_threshold"])
        super().__init__(reward_type=reward_type, **kwargs)
