"""Host-side logic that must work (or fail loudly) without a GPU."""
import numpy as np
import pytest
import torch

from mujoco_panda_pnp_b200 import _lib, engine, synthetic
from mujoco_panda_pnp_b200.distributed import shard_range, summarize_ik
from mujoco_panda_pnp_b200.envs import FrankaRewardModel, FrankaShelfPNPReward
from mujoco_panda_pnp_b200.skills import IKSolver, JacobianIKController

needs_no_gpu = pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")


def test_param_builders():
    p = engine.ik_params()
    assert (p.max_iters, p.pos_thresh, p.damping, p.step_limit, p.kinematics) == (100, 1e-3, 1e-2, 0.1, 0)
    r = engine.reward_params("sparse", n_tasks=2)
    assert r.sparse == 1 and r.n_tasks == 2 and r.distance_threshold == 0.05 and r.high_pick_z == 0.35
    with pytest.raises(ValueError):
        engine.reward_params("shaped")
    with pytest.raises(KeyError):
        engine.ik_params(kinematics="fast")


def test_alias_and_reference_attribute_names():
    assert IKSolver is JacobianIKController
    env = FrankaShelfPNPReward("dense")
    assert env.distance_threshold == 0.05 and env.task_sequence == ["cube1", "cube2", "cube3"]
    assert env.high_pick_z == 0.35 and env.current_task_index == 0
    np.testing.assert_array_equal(env.VERTICAL_QUAT, [1, 0, 0, 0])
    assert env.HORIZONTAL_QUAT[0] == 0.7071067811865476


@needs_no_gpu
def test_product_fails_loudly_without_gpu(kin_model):
    """No CPU fallback: constructing / calling the operators without CUDA raises."""
    from mujoco_panda_pnp_b200 import KinematicData

    with pytest.raises(_lib.PnpLibraryError):
        JacobianIKController(kin_model, KinematicData(kin_model))
    env = FrankaRewardModel()
    with pytest.raises(_lib.PnpLibraryError):
        env.compute_reward(np.zeros(3), np.zeros(3), {})
    with pytest.raises(_lib.PnpLibraryError):
        engine.set_tree(__import__("mujoco_panda_pnp_b200").KinematicTree.from_mjcf())


def test_reward_info_handling():
    env = FrankaRewardModel()
    with pytest.raises(ValueError, match="batched compute_reward needs"):
        env._state_from_info({}, 4)
    with pytest.raises(ValueError, match="one dict per row"):
        env._state_from_info([{}], 2)
    st = env._state_from_info(None, None)
    assert set(st) == {"ee_pos", "ee_quat", "fingers_width", "task_index"}
    infos = [dict(ee_pos=np.ones(3) * i, ee_quat=[1, 0, 0, 0], fingers_width=0.01 * i, task_index=i) for i in range(3)]
    st = env._state_from_info(infos, 3)
    assert st["ee_pos"].shape == (3, 3) and st["task_index"].tolist() == [0, 1, 2]
    with pytest.raises(ValueError, match="same shape"):
        env.compute_reward(np.zeros((2, 3)), np.zeros((3, 3)), {})
    with pytest.raises(ValueError, match=r"\(3,\) or \(N, 3\)"):
        env.compute_reward(np.zeros(4), np.zeros(4), {})


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 4096, (1 << 24) + 3):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)
    s = summarize_ik(torch.tensor([100, 90, 90, 1500]))
    assert s["success_rate"] == 0.9 and s["mean_iterations"] == 15.0


def test_synthetic_generators_are_deterministic_and_stratified():
    a = synthetic.reward_rows(4096, seed=0, dtype=torch.float64)
    b = synthetic.reward_rows(4096, seed=0, dtype=torch.float64)
    for k in a:
        assert torch.equal(a[k], b[k])
    assert a["achieved_goal"].shape == (4096, 3) and a["ee_quat"].shape == (4096, 4)
    assert a["task_index"].dtype == torch.int32 and int(a["task_index"].max()) == 2
    np.testing.assert_allclose(a["ee_quat"].norm(dim=1).numpy(), 1.0, atol=1e-12)
    d_place = (a["achieved_goal"] - a["desired_goal"]).norm(dim=1)
    assert ((d_place - 0.05).abs() < 1e-6).sum() >= 256  # adversarial tail
    assert (d_place < 0.05).sum() > 100
    q = synthetic.random_joint_configs(100, [-1] * 7, [1] * 7, seed=1)
    assert q.shape == (100, 7) and float(q.abs().max()) <= 1.0
    w = synthetic.waypoint_envs(10, seed=2)
    assert w["q_start"].shape == (10, 7) and w["goal"].shape == (10, 3)


def test_her_future_indices_stay_inside_their_episode():
    """synthetic.her_future_indices('future'): index -1 (keep) or a transition of the same episode, not before the row."""
    import torch
    from mujoco_panda_pnp_b200 import synthetic

    for n, T in ((10_000, 300), (1000, 50), (7, 300)):
        fut = synthetic.her_future_indices(n, T, seed=3, strategy="future").long()
        idx = torch.arange(n)
        rel = fut >= 0
        assert 0.1 < float((~rel).float().mean()) < 0.35 or n < 100
        assert bool((fut[rel] >= idx[rel]).all()) and bool((fut[rel] < n).all())
        assert bool(((fut[rel] // T) == (idx[rel] // T)).all())
        assert torch.equal(fut, synthetic.her_future_indices(n, T, seed=3, strategy="future").long())
    uni = synthetic.her_future_indices(10_000, 300, seed=3, strategy="uniform").long()
    assert int(uni.max()) < 10_000 and int(uni.min()) == -1
    assert float(((uni // 300) == (torch.arange(10_000) // 300))[uni >= 0].float().mean()) < 0.1
