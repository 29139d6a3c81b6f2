"""GPU parity: FK + Jacobian kernels vs the oracle (mj_kinematics / mj_jacSite / mju_mat2Quat).
north_star tolerance: 1e-5 m / 1e-5 rad for the FP32 path (stated per assert)."""
import numpy as np
import pytest
import torch

from mujoco_panda_pnp_b200 import KinematicTree, engine, synthetic
from oracle import c_oracle

pytestmark = pytest.mark.gpu

FK_TOL_M = 1e-5  # north_star: FK agrees with mj_kinematics to 1e-5 m
JAC_TOL = 1e-5  # and mj_jacSite to 1e-5 (m/rad for jacp, unitless for jacr)


@pytest.fixture(scope="module")
def tree(cuda_lib):
    t = KinematicTree.from_mjcf()
    assert engine.set_tree(t) is True  # packaged tree -> specialised kernels
    return t


def _quat_close(a, b, tol):
    return np.minimum(np.abs(a - b).max(axis=1), np.abs(a + b).max(axis=1)).max() < tol


@pytest.mark.parametrize("kin", ["specialized", "generic"])
@pytest.mark.parametrize("dtype,tol", [(torch.float32, None), (torch.float64, 1e-12)])
def test_golden_poses(tree, golden_fk, kin, dtype, tol):
    q = torch.tensor(golden_fk["q"], dtype=dtype, device="cuda")
    pos, quat, jac = engine.fk_jac(q, kinematics=kin)
    pos, quat, jac = (x.double().cpu().numpy() for x in (pos, quat, jac))
    np.testing.assert_allclose(pos, golden_fk["pos"], atol=tol or FK_TOL_M)
    np.testing.assert_allclose(jac, golden_fk["jac"], atol=tol or JAC_TOL)
    assert _quat_close(quat, golden_fk["quat"], tol or 1e-5)
    # real-MuJoCo pin (execute_pnp.py:38), printed to 8 decimals
    assert np.abs(pos[0] - golden_fk["home_wpt"]).max() < (5e-9 if dtype == torch.float64 else 1e-6)


@pytest.mark.parametrize("kin", ["specialized", "generic"])
def test_one_million_random_configurations_fp32(tree, oracle_chain, kin):
    n = 1 << 20
    q = synthetic.random_joint_configs(n, tree.lower, tree.upper, seed=42, device="cuda")
    pos, quat, jac = engine.fk_jac(q, kinematics=kin)
    qh = q.double().cpu().numpy()
    want_pos, want_mat, want_jac = c_oracle.fk_jac(oracle_chain, qh, nthreads=8)
    assert np.abs(pos.double().cpu().numpy() - want_pos).max() < FK_TOL_M
    assert np.abs(jac.double().cpu().numpy() - want_jac).max() < JAC_TOL
    # quaternion -> rotation matrix must match the oracle's site_xmat (1e-5 rad)
    qu = quat.double().cpu().numpy()
    w, x, y, z = qu.T
    R = np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y),
                  2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x),
                  2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)], axis=1).reshape(n, 3, 3)
    assert np.abs(R - want_mat).max() < 1e-5
    np.testing.assert_allclose(np.linalg.norm(qu, axis=1), 1.0, atol=1e-6)


def test_fp64_specialised_equals_generic_and_oracle(tree, oracle_chain):
    n = 1 << 16
    q = synthetic.random_joint_configs(n, tree.lower, tree.upper, seed=7, device="cuda", dtype=torch.float64)
    ps, _, js = engine.fk_jac(q, kinematics="specialized")
    pg, _, jg = engine.fk_jac(q, kinematics="generic")
    want_pos, _, want_jac = c_oracle.fk_jac(oracle_chain, q.cpu().numpy(), nthreads=8)
    for p, j in ((ps, js), (pg, jg)):
        assert np.abs(p.cpu().numpy() - want_pos).max() < 1e-13
        assert np.abs(j.cpu().numpy() - want_jac).max() < 1e-13


def test_generic_path_handles_an_arbitrary_tree(cuda_lib, tmp_path):
    """Upload a non-Panda chain: the library must fall back to the generic kernels and agree
    with mj_kinematics on that model."""
    from test_tree import _GENERAL_MJCF
    from mujoco_panda_pnp_b200 import KinematicModel
    from oracle import ik_oracle, mj_oracle

    p = tmp_path / "general.xml"
    p.write_text(_GENERAL_MJCF)
    model = KinematicModel.from_xml_path(str(p))
    t = KinematicTree.from_mjmodel(model)
    try:
        assert engine.set_tree(t) is False
        with pytest.raises(ValueError, match="differs from the build-time"):
            engine.fk_jac(torch.zeros((1, 7), device="cuda"), kinematics="specialized")
        rng = np.random.default_rng(9)
        q = rng.uniform(-2.5, 2.5, (64, 7))
        pos, quat, jac = engine.fk_jac(torch.tensor(q, device="cuda"))
        data = mj_oracle.MjData(model)
        for k in range(len(q)):
            wp, wm, wq, wj = ik_oracle.fk_site(model, data, q[k])
            np.testing.assert_allclose(pos[k].cpu().numpy(), wp, atol=1e-12)
            np.testing.assert_allclose(jac[k].cpu().numpy(), wj, atol=1e-12)
            qk = quat[k].cpu().numpy()
            assert min(np.abs(qk - wq).max(), np.abs(qk + wq).max()) < 1e-12
    finally:
        engine.set_tree(KinematicTree.from_mjcf())


def test_empty_and_validation(tree):
    pos, quat, jac = engine.fk_jac(torch.empty((0, 7), device="cuda"))
    assert pos.shape == (0, 3) and quat.shape == (0, 4) and jac.shape == (0, 6, 7)
    with pytest.raises(ValueError):
        engine.fk_jac(torch.zeros((4, 6), device="cuda"))
    with pytest.raises(ValueError):
        engine.fk_jac(torch.zeros((4, 7)))


@pytest.mark.parametrize("kin", ["specialized", "generic"])
def test_staged_output_kernel_equals_per_lane_kernel(cuda_lib, kin):
    """FP32 batches of >= 128 configurations run fk_jac_bulk_kernel (outputs staged in shared memory and
    sent out with cp.async.bulk stores) plus fk_jac_kernel on the ragged tail; smaller batches run
    fk_jac_kernel alone.  Same arithmetic: chunks of 100 must reproduce the big batch."""
    from mujoco_panda_pnp_b200 import KinematicTree, engine, synthetic

    tree = KinematicTree.from_mjcf()
    engine.set_tree(tree)
    n = 128 * 37 + 77
    q = synthetic.random_joint_configs(n, tree.lower, tree.upper, seed=3, device="cuda")
    for kw in (dict(), dict(want_jac=False), dict(want_quat=False, want_jac=False)):
        big = engine.fk_jac(q, kinematics=kin, **kw)
        for off in (0, 1000, 128 * 37 - 50):
            small = engine.fk_jac(q[off:off + 100].contiguous(), kinematics=kin, **kw)
            for a, b in zip(big, small):
                if a is not None:
                    assert float((a[off:off + 100] - b).abs().max()) <= 1e-6
