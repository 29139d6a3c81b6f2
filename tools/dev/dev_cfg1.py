"""Developer scratch: single-query solve() latency (cfg1)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch
from mujoco_panda_pnp_b200 import KinematicData, KinematicModel, engine, synthetic
from mujoco_panda_pnp_b200.skills import JacobianIKController
kmodel = KinematicModel.from_xml_path(os.path.join(ROOT, "mujoco_panda_pnp_b200", "assets", "panda_shelf_kinematic.xml"))
ctl = JacobianIKController(kmodel, KinematicData(kmodel))
NEUTRAL = np.array(synthetic.NEUTRAL_Q)
grasp = [np.array(t) for t in [(1.415, 0, 0.73), (1.415, 0, 1.03), (1.415, 0, 0.43)]]
for t_ in grasp: ctl.solve(t_, NEUTRAL)
t0 = time.perf_counter(); reps = 200; its = []
for _ in range(reps):
    for t_ in grasp: its.append(ctl.solve(t_, NEUTRAL).iterations)
print(f"solve(): {(time.perf_counter() - t0) / (3 * reps) * 1e6:.1f} us per call, iterations {its[:3]}")
p = engine.ik_params(); t32 = grasp[0].astype(np.float32); q32 = NEUTRAL.astype(np.float32); out = np.empty(12, np.float32)
t0 = time.perf_counter()
for _ in range(1000): engine.ik_solve_one_host(t32, q32, p, out)
print(f"engine.ik_solve_one_host: {(time.perf_counter() - t0) / 1000 * 1e6:.1f} us per call")
