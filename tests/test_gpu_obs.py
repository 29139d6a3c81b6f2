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


@pytest.mark.parametrize("kin", ["specialized", "generic"])
def test_bulk_copy_kernel_equals_per_lane_kernel(cuda_lib, kin):
    """FP32 full tiles of aligned arrays run get_obs_bulk_kernel (all data movement on cp.async.bulk);
    unaligned views and tails run get_obs_kernel.  Same row arithmetic: the rows must agree (bitwise up
    to the compiler's FMA contraction, checked at 2 ulp of the value range) for every size/goal mode."""
    engine.set_tree(KinematicTree.from_mjcf())
    gen = torch.Generator(device="cuda")
    gen.manual_seed(11)
    for n in (127, 128, 129, 1000, 128 * 700 + 77):
        rnd = lambda *sh: torch.randn(sh, generator=gen, device="cuda")  # noqa: E731
        # allocate one row more and slice from row 1: 28/8/12/24-byte row strides are then not 16-byte
        # aligned, which forces the per-lane kernel; the contiguous clones are aligned (bulk kernel)
        big = [rnd(n + 1, 7), rnd(n + 1, 7), rnd(n + 1, 2).abs() * 0.02, rnd(n + 1, 3), rnd(n + 1, 4), rnd(n + 1, 6), rnd(n + 1, 3)]
        views = [t[1:] for t in big]
        assert views[0].data_ptr() % 16 != 0
        aligned = [v.clone() for v in views]
        assert all(t.data_ptr() % 16 == 0 for t in aligned)
        for goal_mode in ("per_env", "broadcast"):
            g_v = views[6] if goal_mode == "per_env" else torch.tensor([1.0, -0.1, 0.3], device="cuda")
            g_a = aligned[6] if goal_mode == "per_env" else g_v
            a = engine.get_obs(*aligned[:6], g_a, kinematics=kin)
            b = engine.get_obs(*views[:6], g_v, kinematics=kin)
            assert a.shape == (n, 25)
            assert float((a - b).abs().max()) <= 1e-6, (n, goal_mode)
            assert torch.equal(a[:, [6, 7, 8, 9, 19, 20, 21, 22, 23, 24]], b[:, [6, 7, 8, 9, 19, 20, 21, 22, 23, 24]])
