"""Developer scratch: reward kernel timing (PNP_RW_BLOCKS = resident-block cap used for the grid size)."""
import os, sys, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from mujoco_panda_pnp_b200 import engine, synthetic
dev = torch.device("cuda")
n = 1 << 24
rows = synthetic.reward_rows(n, seed=0, device=dev, dtype=torch.float32)
args = [rows[k] for k in ("achieved_goal", "desired_goal", "ee_pos", "ee_quat", "fingers_width", "task_index")]
out = torch.empty(n, dtype=torch.float32, device=dev)
p = engine.reward_params("dense")
f = lambda: engine.reward(*args, p, want_success=False, out=out)
for _ in range(3): f()
torch.cuda.synchronize(); ts = []
for _ in range(20):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); f(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
ms = statistics.median(ts)
print(f"reward PNP_RW_BLOCKS={os.environ.get('PNP_RW_BLOCKS', 'default')}: {ms:.4f} ms -> {n / ms / 1e6:.2f} G rows/s, {64.0 * n / ms / 1e6:.0f} GB/s")
