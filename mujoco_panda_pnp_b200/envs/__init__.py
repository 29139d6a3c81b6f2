from .panda_env import HORIZONTAL_QUAT, VERTICAL_QUAT, FrankaRewardModel  # noqa: F401
from .shelf_pnp import SHELF_PNP_CONFIG, FrankaShelfPNPReward  # noqa: F401
