"""Developer scratch: time get_obs (bulk vs per-lane kernel via PNP_OBS_BULK)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from mujoco_panda_pnp_b200 import engine, KinematicTree
engine.set_tree(KinematicTree.from_mjcf())
dev = torch.device("cuda")
n_o = 1 << 22
go = torch.Generator(device=dev); go.manual_seed(3)
rnd = lambda *sh: torch.randn(sh, generator=go, device=dev)
o_args = [rnd(n_o, 7), rnd(n_o, 7), rnd(n_o, 2).abs() * 0.02, rnd(n_o, 3), rnd(n_o, 4), rnd(n_o, 6), rnd(n_o, 3)]
for _ in range(3): engine.get_obs(*o_args)
torch.cuda.synchronize(); ts = []
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); engine.get_obs(*o_args); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
ms = sorted(ts)[len(ts) // 2]
print(f"get_obs PNP_OBS_BULK={os.environ.get('PNP_OBS_BULK', '1')}: {ms:.4f} ms -> {n_o / ms / 1e6:.2f} G envs/s, {228.0 * n_o / ms / 1e6:.0f} GB/s algorithmic")
