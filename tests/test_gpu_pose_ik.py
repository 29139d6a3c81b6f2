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


def test_pose_pair_kernel_follows_the_lane_kernel_and_reaches_the_poses(cuda_lib, oracle_chain):
    """ik_pose_solve_v_kernel<F2> (two pose queries per lane, packed arithmetic, different summation order in the 6x6
    solve) against the one-query-per-lane kernel on 400 k reachable poses, broadcast and per-query q_init, with
    unreachable poses mixed in: same convergence flags and iteration counts up to a few threshold flips, joint solutions
    within 1e-3 rad, and the reached poses - under the oracle's FP64 FK - inside the thresholds."""
    from oracle import c_oracle

    tree = KinematicTree.from_mjcf()
    engine.set_tree(tree)
    n = 400_003
    neutral = torch.tensor(NEUTRAL, dtype=torch.float32, device="cuda")
    qg = synthetic.reachable_move_envs(n, tree.lower, tree.upper, seed=21, device="cuda", spread=0.5)["q_goal"]
    pos, quat, _ = engine.fk_jac(qg, want_jac=False)
    pos[::101] = torch.tensor([2.5, 0.0, 0.5], device="cuda")  # unreachable: runs out of iterations
    q0 = (neutral + 0.1 * torch.randn((n, 7), device="cuda")).contiguous()
    for qi in (neutral, q0):
        ca = torch.zeros(4, dtype=torch.int64, device="cuda")
        cb = torch.zeros(4, dtype=torch.int64, device="cuda")
        a = engine.ik_pose_solve(pos, quat, qi, engine.ik_params(kinematics="spec_lane"), counters=ca)
        b = engine.ik_pose_solve(pos, quat, qi, engine.ik_params(kinematics="spec_pair"), counters=cb)
        assert int(cb[0]) == n and int(cb[1]) == int(b["converged"].sum()) and int(cb[3]) == int(b["iterations"].long().sum())
        same_conv = (a["converged"] == b["converged"])
        same_it = (a["iterations"] == b["iterations"])
        assert float(same_conv.float().mean()) > 0.9995 and float(same_it.float().mean()) > 0.995
        both = a["converged"] & b["converged"] & same_it
        assert float((a["q"][both] - b["q"][both]).abs().max()) < 1e-3
        assert bool((b["iterations"][::101] == 100).all()) and not bool(b["converged"][::101].any())
        # reference FK of a sample of the pair kernel's solutions
        m = 4096
        conv = b["converged"][:m].cpu().numpy()
        ee, mat, _ = c_oracle.fk_jac(oracle_chain, b["q"][:m].double().cpu().numpy(), nthreads=8)
        tp, tq = pos[:m].double().cpu().numpy(), quat[:m].double().cpu().numpy()
        assert np.linalg.norm(ee[conv] - tp[conv], axis=1).max() < 1e-3 + 1e-5
        w, x, y, z = (tq / np.linalg.norm(tq, axis=1, keepdims=True)).T
        tm = np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y),
                       2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x),
                       2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)], axis=1).reshape(-1, 3, 3)
        ang = np.arccos(np.clip((np.einsum("nij,nij->n", mat, tm) - 1) * 0.5, -1, 1))
        assert ang[conv].max() < 1e-2 + 1e-4
    # a q_init outside the joint limits whose pose is already reached: returned untouched after one pass
    q_out = q0[:70_000].clone()
    q_out[::7, 0] = 3.1
    p1, u1, _ = engine.fk_jac(q_out, want_jac=False)
    r = engine.ik_pose_solve(p1, u1, q_out, engine.ik_params(kinematics="spec_pair"))
    sel = torch.arange(0, 70_000, 7, device="cuda")
    assert bool((r["iterations"][sel] == 1).all()) and torch.equal(r["q"][sel], q_out[sel])
