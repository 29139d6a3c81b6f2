"""GPU parity: batched _get_obs (panda_env.py:279-301) vs the reference-generated golden rows."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from mujoco_panda_pnp_b200 import KinematicTree, engine
from mujoco_panda_pnp_b200.envs import FrankaShelfPNPReward

pytestmark = pytest.mark.gpu

KEYS = ("q_arm", "qvel_arm", "fingers", "obj_pos", "obj_quat", "obj_vel", "goal")


def _angle_diff(a, b):
    d = np.abs(a - b)
    return np.minimum(d, 2 * np.pi - d)


@pytest.mark.parametrize("kin", ["specialized", "generic"])
@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-11), (torch.float32, 1e-5)])
def test_get_obs_matches_reference_rows(cuda_lib, kin, dtype, tol):
    """FP32 tolerance 1e-5 m / 1e-5 rad (north_star FK/Jacobian bar); FP64 1e-11."""
    g = np.load(os.path.join(GOLDEN, "obs_reference_golden.npz"))
    engine.set_tree(KinematicTree.from_mjcf())
    args = [torch.tensor(g[k], dtype=dtype, device="cuda") for k in KEYS]
    rows = engine.get_obs(*args, dt=float(g["dt"]), kinematics=kin).double().cpu().numpy()
    assert rows.shape == (256, 25)
    want = np.concatenate([g["observation"], g["achieved_goal"], g["desired_goal"]], axis=1)
    lin = [i for i in range(25) if i not in (10, 11, 12)]
    assert np.abs(rows[:, lin] - want[:, lin]).max() < tol
    # euler angles: compare modulo 2*pi; the exact-gimbal rows (first two) are ill-conditioned in FP32
    sl = slice(0, None) if dtype == torch.float64 else slice(4, None)
    assert _angle_diff(rows[sl, 10:13], want[sl, 10:13]).max() < (tol if dtype == torch.float64 else 2e-4)


def test_env_level_get_obs_dict(cuda_lib):
    g = np.load(os.path.join(GOLDEN, "obs_reference_golden.npz"))
    env = FrankaShelfPNPReward("dense")
    out = env._get_obs({k: g[k] for k in KEYS}, precision="fp64")
    assert set(out) == {"observation", "achieved_goal", "desired_goal"} and out["observation"].shape == (256, 19)
    np.testing.assert_allclose(out["observation"][:, :10], g["observation"][:, :10], atol=1e-11)
    np.testing.assert_array_equal(out["achieved_goal"], g["achieved_goal"])
    np.testing.assert_array_equal(out["desired_goal"], g["desired_goal"])
    # broadcast goal + CUDA tensors in -> CUDA tensors out
    st = {k: torch.tensor(g[k], dtype=torch.float32, device="cuda") for k in KEYS}
    st["goal"] = torch.tensor([1.0, -0.1, 0.3], device="cuda")
    out = env._get_obs(st)
    assert out["observation"].is_cuda and torch.all(out["desired_goal"] == st["goal"])
    with pytest.raises(ValueError):
        env._get_obs({"q_arm": g["q_arm"]})
    assert engine.get_obs(*[torch.empty((0, t.shape[1]), device="cuda") for t in
                            [torch.tensor(g[k]) for k in KEYS]]).shape == (0, 25)
