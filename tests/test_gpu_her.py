"""GPU parity: HER relabel + reward + VecNormalize kernel (bulk-async-copy staged) vs the oracle."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from mujoco_panda_pnp_b200 import engine, synthetic
from oracle import her_oracle

pytestmark = pytest.mark.gpu


def _bits(x):
    return np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)


def _transitions(n, seed, device="cuda"):
    """Stored transitions in the 25-wide layout built from the cfg3 row generator."""
    rows = synthetic.reward_rows(n, seed=seed, device=device, dtype=torch.float32, n_adversarial=min(n // 8, 1 << 12))
    g = torch.Generator(device=device)
    g.manual_seed(seed + 1)
    nxt = torch.randn((n, 25), generator=g, device=device) * 0.3
    nxt[:, 0:3] = rows["ee_pos"]
    nxt[:, 6] = rows["fingers_width"]
    nxt[:, 19:22] = rows["achieved_goal"]
    nxt[:, 7:10] = rows["achieved_goal"]
    nxt[:, 22:25] = rows["desired_goal"]
    obs = nxt + torch.randn((n, 25), generator=g, device=device) * 0.01
    obs[:, 22:25] = rows["desired_goal"]
    fut = torch.randint(0, max(n, 1), (n,), generator=g, device=device, dtype=torch.int32)
    keep = torch.rand((n,), generator=g, device=device) < 0.2   # 20 % keep the real goal
    fut = torch.where(keep, torch.full_like(fut, -1), fut)
    # make a share of the relabelled goals "placed": future = the row itself
    own = (torch.arange(n, device=device) % 16) == 3
    fut = torch.where(own, torch.arange(n, device=device, dtype=torch.int32), fut)
    return obs.contiguous(), nxt.contiguous(), fut.contiguous(), rows["ee_quat"], rows["task_index"]


def _norm_stats():
    with open(os.path.join(GOLDEN, "vecnormalize_stats.json")) as fh:
        st = json.load(fh)["obs_rms"]
    mean = np.concatenate([st["observation"]["mean"], st["achieved_goal"]["mean"], st["desired_goal"]["mean"]])
    var = np.concatenate([st["observation"]["var"], st["achieved_goal"]["var"], st["desired_goal"]["var"]])
    return mean, var


@pytest.mark.parametrize("n", [1, 5, 127, 128, 129, 1000, 4096 + 77])
@pytest.mark.parametrize("rt", ["dense", "sparse"])
def test_relabel_copy_through_is_bit_exact(cuda_lib, n, rt):
    obs, nxt, fut, quat, task = _transitions(n, seed=n)
    cnt = torch.zeros(4, dtype=torch.int64, device="cuda")
    o, x, r, s = engine.her_relabel(obs, nxt, fut, quat, task, engine.reward_params(rt), counters=cnt)
    wo, wx, wr, ws = her_oracle.relabel(obs.cpu().numpy(), nxt.cpu().numpy(), fut.cpu().numpy(), quat.cpu().numpy(),
                                        task.cpu().numpy(), reward_type=rt)
    np.testing.assert_array_equal(_bits(o.cpu().numpy()), _bits(wo))
    np.testing.assert_array_equal(_bits(x.cpu().numpy()), _bits(wx))
    np.testing.assert_array_equal(_bits(r.cpu().numpy()), _bits(wr))
    np.testing.assert_array_equal(s.cpu().numpy(), ws)
    c = cnt.cpu().numpy()
    assert c[0] == n and c[1] == int(ws.sum())


def test_relabel_with_vecnormalize_and_unaligned_views(cuda_lib):
    mean, var = _norm_stats()
    norm = engine.normalize_params(mean, var)  # the reference pickle's statistics, clip 10, eps 1e-8
    n = 3000
    obs, nxt, fut, quat, task = _transitions(n + 1, seed=5)
    for off in (0, 1):  # off=1: row views start 100 B into the buffer -> not 16-byte aligned -> plain path
        sl = slice(off, off + n)
        f = torch.clamp(fut[sl] - off, min=-1).contiguous()
        f = torch.where(f >= n, torch.full_like(f, -1), f)
        o, x, r, s = engine.her_relabel(obs[sl], nxt[sl], f, quat[sl].contiguous(), task[sl].contiguous(),
                                        engine.reward_params("dense"), norm=norm)
        wo, wx, wr, ws = her_oracle.relabel(obs[sl].cpu().numpy(), nxt[sl].cpu().numpy(), f.cpu().numpy(),
                                            quat[sl].cpu().numpy(), task[sl].cpu().numpy(), mean=mean, var=var)
        # FP32 ((x - mean_hi) - mean_lo) * inv_std vs the oracle's float64 division: a few ulp of
        # values clipped to +-10
        np.testing.assert_allclose(o.cpu().numpy(), wo, atol=5e-6, rtol=2e-6)
        np.testing.assert_allclose(x.cpu().numpy(), wx, atol=5e-6, rtol=2e-6)
        assert float(o.abs().max()) <= 10.0 and float((o.abs() == 10.0).float().mean()) > 0.0  # clip is active
        np.testing.assert_array_equal(_bits(r.cpu().numpy()), _bits(wr))
        np.testing.assert_array_equal(s.cpu().numpy(), ws)


def test_relabel_rejects_aliasing_and_bad_shapes(cuda_lib):
    obs, nxt, fut, quat, task = _transitions(256, seed=1)
    with pytest.raises(ValueError, match="alias"):
        engine.her_relabel(obs, nxt, fut, quat, task, engine.reward_params(), out_obs=obs)
    with pytest.raises(ValueError):
        engine.her_relabel(obs[:, :24].contiguous(), nxt, fut, quat, task, engine.reward_params())
    with pytest.raises(ValueError):
        engine.normalize_params(np.zeros(19), np.ones(19))
    o, x, r, s = engine.her_relabel(obs[:0], nxt[:0], fut[:0], quat[:0], task[:0], engine.reward_params())
    assert o.shape == (0, 25) and r.shape == (0,)


def test_relabel_4m_transitions_properties(cuda_lib):
    """2^22 transitions: relabelled rows equal the gathered goals, rewards match the plain reward
    kernel bit for bit (two independent code paths over the same arithmetic)."""
    n = 1 << 22
    obs, nxt, fut, quat, task = _transitions(n, seed=9)
    o, x, r, s = engine.her_relabel(obs, nxt, fut, quat, task, engine.reward_params("dense"))
    goal = torch.where((fut >= 0)[:, None], nxt[fut.clamp(min=0).long(), 19:22], nxt[:, 22:25])
    assert torch.equal(o[:, 22:25], goal) and torch.equal(x[:, 22:25], goal)
    assert torch.equal(o[:, :22], obs[:, :22]) and torch.equal(x[:, :22], nxt[:, :22])
    r2, s2 = engine.reward(nxt[:, 19:22].contiguous(), goal.contiguous(), nxt[:, 0:3].contiguous(), quat,
                           nxt[:, 6].contiguous(), task, engine.reward_params("dense"))
    assert torch.equal(r.view(torch.int32), r2.view(torch.int32)) and torch.equal(s, s2)
    assert int(s.sum()) > n // 32  # the "own row" relabels are placed


@pytest.mark.parametrize("n", [5, 129, 4096 + 77, 1 << 20])
def test_relabel_from_goal_table_is_identical(cuda_lib, n):
    """pnp_her_relabel_table_f32: future goals gathered from a separate [N,3] table (what a replay buffer
    keeps as next_observations["achieved_goal"]) instead of out of the 100-byte rows - identical outputs."""
    obs, nxt, fut, quat, task = _transitions(n, seed=3 + n % 11)
    mean, var = _norm_stats()
    table = nxt[:, 19:22].contiguous()
    for norm in (None, engine.normalize_params(mean, var)):
        a = engine.her_relabel(obs, nxt, fut, quat, task, engine.reward_params("dense"), norm=norm)
        b = engine.her_relabel(obs, nxt, fut, quat, task, engine.reward_params("dense"), norm=norm, future_ag=table)
        for u, v in zip(a, b):
            assert torch.equal(u.view(torch.int32), v.view(torch.int32))
    with pytest.raises(ValueError):
        engine.her_relabel(obs, nxt, fut, quat, task, engine.reward_params("dense"), future_ag=table[:-1])


@pytest.mark.parametrize("episode_len", [50, 300])
def test_relabel_future_strategy_episode_indices(cuda_lib, episode_len):
    """The bench's headline HER workload: goals from a later transition of the same episode (synthetic.her_future_indices).
    Bit-exact against the oracle on a sample that spans several tiles and a ragged last episode."""
    n = 5 * 300 + 77
    obs, nxt, _, quat, task = _transitions(n, seed=21)
    fut = synthetic.her_future_indices(n, episode_len, seed=4, device="cuda", strategy="future")
    o, x, r, s = engine.her_relabel(obs, nxt, fut, quat, task, engine.reward_params("dense"))
    wo, wx, wr, ws = her_oracle.relabel(obs.cpu().numpy(), nxt.cpu().numpy(), fut.cpu().numpy(), quat.cpu().numpy(),
                                        task.cpu().numpy(), reward_type="dense")
    np.testing.assert_array_equal(_bits(o.cpu().numpy()), _bits(wo))
    np.testing.assert_array_equal(_bits(x.cpu().numpy()), _bits(wx))
    np.testing.assert_array_equal(_bits(r.cpu().numpy()), _bits(wr))
    np.testing.assert_array_equal(s.cpu().numpy(), ws)
