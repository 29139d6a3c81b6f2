"""GPU: pose-mode (6-row, 6x6 solve) IK extension vs its FP64 NumPy statement (oracle/pose_ik_oracle.py).
No reference behaviour exists for this entry point (panda_env.py:399-409 imports a missing name)."""
import numpy as np
import pytest
import torch

from conftest import NEUTRAL
from mujoco_panda_pnp_b200 import KinematicData, KinematicTree, engine, synthetic
from mujoco_panda_pnp_b200.skills import JacobianIKController, solve_ik
from oracle import ik_oracle, mj_oracle, pose_ik_oracle

pytestmark = pytest.mark.gpu


def _targets(oracle_model, n, seed, spread=0.5):
    rng = np.random.default_rng(seed)
    lo, hi = oracle_model.jnt_range[:7, 0], oracle_model.jnt_range[:7, 1]
    qs = np.clip(NEUTRAL + rng.uniform(-spread, spread, (n, 7)), lo, hi)
    d = mj_oracle.MjData(oracle_model)
    pos, quat = [], []
    for q in qs:
        p, _, qu, _ = ik_oracle.fk_site(oracle_model, d, q)
        pos.append(p), quat.append(qu)
    return np.array(pos), np.array(quat)


@pytest.mark.parametrize("kin", ["specialized", "generic"])
def test_fp64_pose_ik_is_iteration_exact_vs_statement(cuda_lib, oracle_model, kin_model, kin):
    pos, quat = _targets(oracle_model, 48, seed=1)
    ctl = JacobianIKController(kin_model, KinematicData(kin_model), precision="fp64", kinematics=kin)
    out = ctl.solve_pose(pos, quat, NEUTRAL)
    d = mj_oracle.MjData(oracle_model)
    n_conv = 0
    for k in range(len(pos)):
        w = pose_ik_oracle.solve_pose(oracle_model, d, pos[k], quat[k], NEUTRAL)
        assert int(out["iterations"][k]) == w["iterations"], k
        assert bool(out["converged"][k]) == w["converged"] and bool(out["success"][k]) == w["success"]
        tol = 1e-8 if w["converged"] else 1e-5
        np.testing.assert_allclose(out["q"][k].cpu().numpy(), w["q"], atol=tol)
        np.testing.assert_allclose(out["final_pos"][k].cpu().numpy(), w["final_pos"], atol=tol)
        assert abs(float(out["rot_error"][k]) - w["rot_error"]) < tol
        n_conv += w["converged"]
    assert n_conv >= 40


def test_fp32_pose_ik_reaches_the_pose(cuda_lib, oracle_model, oracle_chain, kin_model):
    from oracle import c_oracle

    pos, quat = _targets(oracle_model, 2048 // 8, seed=2)
    ctl = JacobianIKController(kin_model, KinematicData(kin_model))
    out = ctl.solve_pose(pos, quat, NEUTRAL)
    conv = out["converged"].cpu().numpy()
    assert conv.mean() > 0.9
    q = out["q"].double().cpu().numpy()
    ee, mat, _ = c_oracle.fk_jac(oracle_chain, q)
    assert np.linalg.norm(ee[conv] - pos[conv], axis=1).max() < 1e-3 + 1e-5
    # orientation: angle between the reached and the requested frame below rot_thresh (+ FP32 slack)
    for k in np.nonzero(conv)[0][:64]:
        qc = np.empty(4)
        mj_oracle.mju_mat2Quat(qc, mat[k].reshape(9))
        ang = 2 * np.arccos(min(1.0, abs(float(qc @ quat[k]))))
        assert ang < 1e-2 + 1e-4
    # the dangling reference API: solve_ik(model, data, site_name, target_pos, target_quat, q_init)
    data = KinematicData(kin_model)
    r = solve_ik(kin_model, data, "ee_center_site", pos[0], quat[0], NEUTRAL)
    assert r.success and r.q.shape == (7,) and np.array_equal(data.qpos[:7], r.q)
    with pytest.raises(ValueError):
        engine.ik_pose_solve(torch.zeros((2, 3), device="cuda"), torch.zeros((3, 4), device="cuda"),
                             torch.zeros(7, device="cuda"), engine.ik_params())
