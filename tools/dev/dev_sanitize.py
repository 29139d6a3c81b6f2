"""Small-size pass over every kernel family, meant to run under compute-sanitizer (memcheck; `race` as argv[1] keeps
only the kernels with shared-memory hand-overs, for racecheck)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch
from mujoco_panda_pnp_b200 import engine, synthetic, KinematicTree
tree = KinematicTree.from_mjcf(); engine.set_tree(tree)
dev = torch.device("cuda")
neutral = torch.tensor(synthetic.NEUTRAL_Q, device=dev)
RACE = len(sys.argv) > 1 and sys.argv[1] == "race"
# round 2: the pair kernel's tail phase (needs a batch that fills the grid: spec_pair at 200 k queries, some of them
# unreachable so that stragglers are parked), compact records, the small-batch latency kernel (AUTO at n <= 128 per SM)
big = synthetic.random_joint_configs(200_003, tree.lower, tree.upper, seed=4, device=dev)
tgb = engine.fk_jac(big, want_quat=False, want_jac=False)[0]
tgb[::61] = torch.tensor([2.5, 0.0, 0.5], device=dev)
for compact in (False, True):
    engine.ik_solve(tgb, neutral, engine.ik_params(kinematics="spec_pair", max_iters=40), compact=compact)
engine.ik_solve(tgb[:4096], neutral, engine.ik_params(max_iters=40))
engine.ik_solve(tgb[:33], neutral, engine.ik_params(max_iters=40), compact=True)
torch.cuda.synchronize()
if RACE:
    print("sanitize pass (race subset) done")
    sys.exit(0)
import numpy as _np
from mujoco_panda_pnp_b200 import KinematicData, KinematicModel
from mujoco_panda_pnp_b200.envs import FrankaShelfPNPReward
from mujoco_panda_pnp_b200.skills import JacobianIKController
_model = KinematicModel.from_xml_path(os.path.join(ROOT, "mujoco_panda_pnp_b200", "assets", "panda_shelf_kinematic.xml"))
JacobianIKController(_model, KinematicData(_model)).solve(_np.array([1.415, 0.0, 0.73]), _np.array(synthetic.NEUTRAL_Q))
FrankaShelfPNPReward("dense").compute_reward(_np.array([1.0, 0.0, 0.3]), _np.array([1.0, 0.1, 0.3]), {})
order = torch.randperm(5000, device=dev).int()
engine.move_plan_order_check(order)
for n in (1, 65, 1000, 70001):
    q = synthetic.random_joint_configs(n, tree.lower, tree.upper, seed=1, device=dev)
    tg = engine.fk_jac(q)[0]
    q0 = (neutral + 0.1 * torch.randn(n, 7, device=dev)).contiguous()
    for kin in ("generic", "spec_lane", "spec_pair"):
        for qi in (neutral, q0):
            for packed in (True, False):
                engine.ik_solve(tg, qi, engine.ik_params(kinematics=kin, max_iters=30), packed=packed)
    engine.ik_solve(tg.double(), neutral.double(), engine.ik_params(max_iters=20))
    w = synthetic.waypoint_envs(n, seed=2, device=dev)
    for kin in ("generic", "spec_lane", "spec_pair"):
        engine.ik_waypoints(w["q_start"], w["goal"], 12, engine.ik_params(kinematics=kin))
    m = min(n, 2000)
    engine.move_ik_plan(w["q_start"][:m].contiguous(), w["goal"][:m].contiguous(), engine.ik_params())
    rnd = lambda *sh: torch.randn(sh, device=dev)
    engine.get_obs(rnd(n, 7), rnd(n, 7), rnd(n, 2), rnd(n, 3), rnd(n, 4), rnd(n, 6), rnd(n, 3))
    engine.get_obs(rnd(n, 7), rnd(n, 7), rnd(n, 2), rnd(n, 3), rnd(n, 4), rnd(n, 6), torch.tensor([1.0, 0.0, 0.3], device=dev))
    rows = synthetic.reward_rows(n, seed=3, device=dev, dtype=torch.float32)
    args = [rows[k] for k in ("achieved_goal", "desired_goal", "ee_pos", "ee_quat", "fingers_width", "task_index")]
    engine.reward(*args, engine.reward_params("dense"))
    nxt = rnd(n, 25); obs = nxt + 0.01
    fut = torch.randint(-1, n, (n,), device=dev, dtype=torch.int32)
    quat = rnd(n, 4); task = torch.randint(0, 3, (n,), device=dev, dtype=torch.int32)
    engine.her_relabel(obs, nxt, fut, quat, task, engine.reward_params("dense"))
    engine.her_relabel(obs, nxt, fut, quat, task, engine.reward_params("dense"), future_ag=nxt[:, 19:22].contiguous())
    pp, pq, _ = engine.fk_jac(q)
    engine.ik_pose_solve(pp, pq, neutral.repeat(n, 1).contiguous(), engine.ik_params(max_iters=20))
torch.cuda.synchronize()
print("sanitize pass done")
