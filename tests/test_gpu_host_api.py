"""GPU: the host-buffer C-ABI operators (the e2e path bench.py times)."""
import numpy as np
import pytest
import torch

from conftest import NEUTRAL
from mujoco_panda_pnp_b200 import KinematicTree, engine, synthetic
from oracle import c_oracle

pytestmark = pytest.mark.gpu


def test_ik_host_pipeline_equals_device_path(cuda_lib, oracle_chain):
    tree = KinematicTree.from_mjcf()
    engine.set_tree(tree)
    n = 50_003
    qstar = synthetic.random_joint_configs(n, tree.lower, tree.upper, seed=3, dtype=torch.float64).numpy()
    targets = c_oracle.fk_jac(oracle_chain, qstar, nthreads=8)[0].astype(np.float32)
    p = engine.ik_params()
    dev = engine.ik_solve(torch.tensor(targets, device="cuda"), torch.tensor(NEUTRAL, dtype=torch.float32, device="cuda"), p)
    for chunk, packed in ((0, True), (1000, True), (1000, False)):  # 1 chunk / 51 chunks over 3 streams
        r = engine.ik_solve_host(targets, NEUTRAL.astype(np.float32), p, chunk_rows=chunk, packed=packed)
        np.testing.assert_array_equal(r["q"], dev.q.cpu().numpy())
        np.testing.assert_array_equal(r["iterations"], dev.iterations.cpu().numpy())
        np.testing.assert_array_equal(r["final_pos"], dev.final_pos.cpu().numpy())
        np.testing.assert_array_equal(r["pos_error"], dev.pos_error.cpu().numpy())
        np.testing.assert_array_equal(r["converged"], dev.converged.cpu().numpy())
        assert r["counters"][0] == n and r["counters"][1] == int(dev.converged.sum())
        assert r["counters"][3] == int(dev.iterations.sum())
    # per-query q_init through the host path + pinned, preallocated outputs
    q0 = np.clip(NEUTRAL + np.random.default_rng(0).uniform(-0.2, 0.2, (n, 7)), tree.lower, tree.upper).astype(np.float32)
    out = dict(q8=torch.empty((n, 8)).pin_memory().numpy(), aux4=torch.empty((n, 4)).pin_memory().numpy())
    r = engine.ik_solve_host(torch.tensor(targets).pin_memory(), q0, p, chunk_rows=7777, out=out)
    assert r["q8"] is out["q8"] and np.shares_memory(r["q"], out["q8"]) and np.shares_memory(r["final_pos"], out["aux4"])
    dev2 = engine.ik_solve(torch.tensor(targets, device="cuda"), torch.tensor(q0, device="cuda"), p)
    np.testing.assert_array_equal(r["q"], dev2.q.cpu().numpy())
    with pytest.raises(ValueError):
        engine.ik_solve_host(targets, np.zeros((3, 7), np.float32), p)


def test_fp32_peak_probe_and_launch_counter(cuda_lib):
    before = cuda_lib.pnp_launch_count()
    tflops, ms = engine.probe_fp32_peak()
    assert 30.0 < tflops < 90.0 and ms > 0  # B200: 148 SMs x 128 lanes x 2 x ~1.9 GHz = 72-74
    assert cuda_lib.pnp_launch_count() == before + 2
