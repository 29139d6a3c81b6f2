"""SURVEY.md Appendix C known-answer vectors, as literals, against the oracle (CPU) and the kernels (GPU).
Printed precision in the survey is 8 significant digits, hence the 1e-7 / 5e-8 tolerances; reward words
are compared bit for bit."""
import numpy as np
import pytest

import appendix_c as C
from oracle import c_oracle, ik_oracle, mj_oracle, reward_oracle


def _bits(x):
    return np.asarray(x, dtype=np.float32).view(np.uint32)


# ------------------------------------------------------------------------------------ CPU: the oracle
def test_oracle_fk_jacobian_known_answers(oracle_model):
    d = mj_oracle.MjData(oracle_model)
    pos, _, quat, jac = ik_oracle.fk_site(oracle_model, d, C.NEUTRAL)
    np.testing.assert_allclose(pos, C.FK_NEUTRAL_POS, atol=5e-9)        # real-MuJoCo value
    np.testing.assert_allclose(quat, C.FK_NEUTRAL_QUAT, atol=5e-9)
    np.testing.assert_allclose(jac[:3], C.FK_NEUTRAL_JP, atol=5e-9)
    np.testing.assert_allclose(jac[3:], C.FK_NEUTRAL_JR, atol=5e-9)
    pos0, _, quat0, _ = ik_oracle.fk_site(oracle_model, d, np.zeros(7))
    np.testing.assert_allclose(pos0, C.FK_ZERO_POS, atol=5e-9)
    np.testing.assert_allclose(quat0, C.FK_ZERO_QUAT, atol=5e-9)


def test_oracle_ik_known_answers(oracle_model, oracle_chain):
    ctl = ik_oracle.JacobianIKController(oracle_model, mj_oracle.MjData(oracle_model))
    for target, kw, iters, q, final_pos, err in C.IK_CASES:
        r = ctl.solve(np.array(target), C.NEUTRAL, **kw)
        assert r.converged and r.success and r.iterations == iters, target
        np.testing.assert_allclose(r.q, q, atol=5e-8)
        np.testing.assert_allclose(r.final_pos, final_pos, atol=5e-8)
        assert abs(r.pos_error - err) < max(5e-8, 1e-4 * err)
        rc = c_oracle.ik_solve(oracle_chain, np.array([target]), C.NEUTRAL, **kw)
        assert int(rc["iterations"][0]) == iters
        np.testing.assert_allclose(rc["q"][0], q, atol=5e-8)


def test_oracle_reward_known_answers(oracle_chain):
    ag, dg, ee, eq, w, idx, dense_bits, sparse_bits = C.reward_arrays()
    dense = reward_oracle.compute_reward_rows(ag, dg, ee, eq, w, idx, reward_type="dense")
    sparse = reward_oracle.compute_reward_rows(ag, dg, ee, eq, w, idx, reward_type="sparse")
    np.testing.assert_array_equal(_bits(dense), dense_bits)
    np.testing.assert_array_equal(_bits(sparse), sparse_bits)            # incl. -0.0 (0x80000000) when placed
    for k, row in enumerate(C.REWARD_ROWS):
        assert abs(float(dense[k]) - row[6]) < 1e-6


# ------------------------------------------------------------------------------------ GPU: the kernels
@pytest.mark.gpu
@pytest.mark.parametrize("kin", ["specialized", "generic"])
def test_kernels_fk_jacobian_known_answers(cuda_lib, kin):
    import torch
    from mujoco_panda_pnp_b200 import KinematicTree, engine

    engine.set_tree(KinematicTree.from_mjcf())
    for dtype, tol in ((torch.float64, 5e-9), (torch.float32, 1e-5)):  # north_star: FK / Jacobian to 1e-5
        q = torch.tensor(np.stack([C.NEUTRAL, np.zeros(7)]), dtype=dtype, device="cuda")
        pos, quat, jac = engine.fk_jac(q, kinematics=kin)
        pos, quat, jac = pos.double().cpu().numpy(), quat.double().cpu().numpy(), jac.double().cpu().numpy()
        np.testing.assert_allclose(pos[0], C.FK_NEUTRAL_POS, atol=tol)
        np.testing.assert_allclose(quat[0], C.FK_NEUTRAL_QUAT, atol=tol)
        np.testing.assert_allclose(jac[0].reshape(6, 7)[:3], C.FK_NEUTRAL_JP, atol=tol)
        np.testing.assert_allclose(jac[0].reshape(6, 7)[3:], C.FK_NEUTRAL_JR, atol=tol)
        np.testing.assert_allclose(pos[1], C.FK_ZERO_POS, atol=tol)
        np.testing.assert_allclose(quat[1], C.FK_ZERO_QUAT, atol=tol)


@pytest.mark.gpu
def test_kernels_ik_known_answers(cuda_lib, kin_model):
    from mujoco_panda_pnp_b200 import KinematicData
    from mujoco_panda_pnp_b200.skills import JacobianIKController

    for precision, tol in (("fp64", 5e-8), ("fp32", 2e-5)):
        ctl = JacobianIKController(kin_model, KinematicData(kin_model), precision=precision)
        for target, kw, iters, q, final_pos, err in C.IK_CASES:
            r = ctl.solve(np.array(target), C.NEUTRAL, **kw)
            assert r.converged and r.success and r.iterations == iters, (precision, target)
            np.testing.assert_allclose(r.q, q, atol=tol)
            np.testing.assert_allclose(r.final_pos, final_pos, atol=tol)


@pytest.mark.gpu
def test_kernels_reward_known_answers(cuda_lib):
    import torch
    from mujoco_panda_pnp_b200 import engine

    for dtype in (torch.float64, torch.float32):
        ag, dg, ee, eq, w, idx, dense_bits, sparse_bits = C.reward_arrays()
        args = [torch.tensor(a, dtype=dtype, device="cuda") for a in (ag, dg, ee, eq, w)] + [torch.tensor(idx, device="cuda")]
        for rt, want in (("dense", dense_bits), ("sparse", sparse_bits)):
            r, _ = engine.reward(*args, engine.reward_params(rt))
            got = r.cpu().numpy().view(np.uint32)
            if dtype == torch.float64:
                np.testing.assert_array_equal(got, want)
            else:
                # FP32 storage rounds the inputs first: same bits as the oracle fed the rounded inputs
                ref = reward_oracle.compute_reward_rows(*[a.astype(np.float32).astype(np.float64) for a in (ag, dg, ee, eq, w)], idx, reward_type=rt)
                np.testing.assert_array_equal(got, _bits(ref))
