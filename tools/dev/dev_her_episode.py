"""Developer scratch: HER relabel timing by where the future goal comes from.
uniform = any row of the buffer (worst case for the gather); episode = a later transition of the same episode
(SB3 'future' strategy; episodes of T consecutive rows); none = future_idx -1 (stored goal kept)."""
import os, sys, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch
from mujoco_panda_pnp_b200 import engine
dev = torch.device("cuda")
n_h = 1 << 23
g = torch.Generator(device=dev); g.manual_seed(7)
h_next = torch.randn((n_h, 25), generator=g, device=dev); h_obs = h_next + 0.01
h_quat = torch.randn((n_h, 4), generator=g, device=dev)
h_task = torch.randint(0, 3, (n_h,), generator=g, device=dev, dtype=torch.int32)
h_o, h_x, h_r = torch.empty_like(h_obs), torch.empty_like(h_next), torch.empty(n_h, device=dev)
nrm = engine.normalize_params(np.zeros(25), np.ones(25))
p = engine.reward_params("dense")
tab = h_next[:, 19:22].contiguous()
idx = torch.arange(n_h, device=dev, dtype=torch.int64)
cases = {"uniform": torch.randint(-1, n_h, (n_h,), generator=g, device=dev, dtype=torch.int32),
         "none": torch.full((n_h,), -1, device=dev, dtype=torch.int32)}
for T in (50, 300):
    t = idx % T
    u = torch.rand(n_h, generator=g, device=dev)
    fut = idx + (u * (T - t).float()).long().clamp_(max=T - 1)  # in [t, T-1] of the same episode
    fut = torch.minimum(fut - 0, (idx - t) + T - 1).clamp_(max=n_h - 1)
    keep = torch.rand(n_h, generator=g, device=dev) < 0.2           # 20 % keep the real goal (n_sampled_goal = 4)
    cases[f"episode{T}"] = torch.where(keep, torch.full_like(fut, -1), fut).to(torch.int32)
for name, h_fut in cases.items():
    for tname, kw in (("rows", {}), ("table", {"future_ag": tab})):
        f = lambda: engine.her_relabel(h_obs, h_next, h_fut, h_quat, h_task, p, norm=nrm, want_success=False, out_obs=h_o, out_next_obs=h_x, out_reward=h_r, **kw)
        for _ in range(3): f()
        torch.cuda.synchronize(); ts = []
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); f(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        ms = statistics.median(ts)
        print(f"her_relabel future={name:10s} gather from {tname:5s}: {ms:.4f} ms -> {n_h / ms / 1e6:.2f} G transitions/s, {464.0 * n_h / ms / 1e6:.0f} GB/s (464 B/transition)", flush=True)
