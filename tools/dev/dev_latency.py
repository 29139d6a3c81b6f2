"""Developer scratch: the two latency cases - cfg2 (4096 cold targets, one launch, CUDA-graph replay so that no Python sits
between the events) and cfg1 (one solve() host to host)."""
import os, sys, time, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch
from mujoco_panda_pnp_b200 import KinematicTree, engine, synthetic
tree = KinematicTree.from_mjcf(); engine.set_tree(tree)
dev = torch.device("cuda")
neutral = torch.tensor(synthetic.NEUTRAL_Q, dtype=torch.float32, device=dev)
q = synthetic.random_joint_configs(4096, tree.lower, tree.upper, seed=1234, device=dev)
tg = engine.fk_jac(q, want_quat=False, want_jac=False)[0]
q8, aux = torch.empty((4096, 8), device=dev), torch.empty((4096, 4), device=dev)
p = engine.ik_params()
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    engine.ik_solve(tg, neutral, p, out_q8=q8, out_aux4=aux)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        engine.ik_solve(tg, neutral, p, out_q8=q8, out_aux4=aux)
    ts = []
    for _ in range(50):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
print(f"cfg2 graph replay: median {statistics.median(ts):.2f} us min {min(ts):.2f} us")
t32 = np.array([1.415, 0, 0.73], np.float32); q32 = np.array(synthetic.NEUTRAL_Q, np.float32); out = np.empty(12, np.float32)
for _ in range(200): engine.ik_solve_one_host(t32, q32, p, out)
t0 = time.perf_counter()
for _ in range(2000): engine.ik_solve_one_host(t32, q32, p, out)
print(f"cfg1 engine.ik_solve_one_host: {(time.perf_counter() - t0) / 2000 * 1e6:.2f} us per call")
