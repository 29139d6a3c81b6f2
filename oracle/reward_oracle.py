"""ORACLE (test infrastructure): FP64 restatement of the reward / success arithmetic.

Follows /root/reference/panda_mujoco_gym/envs/panda_env.py:205-245 (compute_reward),
:303-306 (_is_success), :311-315 (goal_distance), constants :29-30, :38-45, :215 and
/root/reference/panda_mujoco_gym/envs/shelf_pnp.py:17-26.  The reference reads EE position,
finger width and EE quaternion from the live simulator; here they are explicit
arguments.  All arithmetic is Python/NumPy float64 in the reference's operation order, one
final cast to float32.

Pinning: tests/golden/reward_reference_golden.json is produced by oracle/gen_golden.py from the
reference's *own* FrankaEnv.compute_reward/_is_success (imported unmodified with its third-party
imports stubbed); row 0 reproduces the real ``old_reward = -0.053`` (bits 0xbd591687) stored in
scripts/checkpoints/tqc_dense_vecnormalize_200000_steps.pkl.
"""

from __future__ import annotations

import numpy as np


def euler2quat(euler):
    """gymnasium_robotics.utils.rotations.euler2quat (gymnasium-robotics==1.2.2, third-party,
    absent from /root/reference; restated from the published source)."""
    euler = np.asarray(euler, dtype=np.float64)
    ai, aj, ak = euler[..., 2] / 2, -euler[..., 1] / 2, euler[..., 0] / 2
    si, sj, sk = np.sin(ai), np.sin(aj), np.sin(ak)
    ci, cj, ck = np.cos(ai), np.cos(aj), np.cos(ak)
    cc, cs = ci * ck, ci * sk
    sc, ss = si * ck, si * sk
    quat = np.empty(euler.shape[:-1] + (4,), dtype=np.float64)
    quat[..., 0] = cj * cc + sj * ss
    quat[..., 3] = cj * sc - sj * cs
    quat[..., 2] = -(cj * ss + sj * cc)
    quat[..., 1] = cj * cs - sj * sc
    return quat


VERTICAL_QUAT = euler2quat(np.zeros(3))  # panda_env.py:29
HORIZONTAL_QUAT = euler2quat(np.array([-np.pi / 2, 0, 0]))  # panda_env.py:30


def goal_distance(a, b):  # panda_env.py:311-315
    a = np.array(a)
    b = np.array(b)
    return np.linalg.norm(a - b, axis=-1)


def is_success(achieved_goal, desired_goal, distance_threshold=0.05):  # panda_env.py:303-306
    d = float(goal_distance(achieved_goal, desired_goal))
    return np.float32(1.0 if d < distance_threshold else 0.0)


def compute_reward(
    achieved_goal,
    desired_goal,
    ee_pos,
    ee_quat,
    fingers_width,
    task_index,
    *,
    reward_type="dense",
    n_tasks=3,
    initial_object_height=0.001,
    distance_threshold=0.05,
    high_pick_z=0.35,
):
    """Scalar restatement of panda_env.py:205-245 with the hidden state made explicit."""
    achieved_goal = np.asarray(achieved_goal)
    desired_goal = np.asarray(desired_goal)
    d_reach = float(goal_distance(ee_pos, achieved_goal))  # :211
    d_place = float(goal_distance(achieved_goal, desired_goal))  # :212
    ee_width = float(fingers_width)  # :214
    GRIP_WIDTH_THRESH = 0.045  # :215
    gripped = (ee_width < GRIP_WIDTH_THRESH) and (d_reach < 0.05)  # :216
    lifted = gripped and (achieved_goal[2] - initial_object_height > 0.04)  # :219
    placed = d_place < distance_threshold  # :220
    ee_q = np.asarray(ee_quat, dtype=np.float64)  # :222
    need_q = HORIZONTAL_QUAT if achieved_goal[2] > high_pick_z else VERTICAL_QUAT  # :223
    ori_err = float(1.0 - abs(np.dot(ee_q, need_q)))  # :224
    if reward_type == "sparse":  # :227-228
        return np.float32(-float(not placed))
    reward = -0.003  # :231
    reward += -min(d_reach, 0.05)  # :232
    if gripped:  # :234-236
        reward += 2.0
        reward += 1.0 - ori_err
    if lifted:  # :238-239
        reward += 4.0
    if placed:  # :241-242
        reward += 10.0
    reward += 0.5 * (int(task_index) / n_tasks)  # :244
    return np.float32(reward)  # :245


def compute_reward_rows(ag, dg, ee_pos, ee_quat, width, task_index, **kw):
    """Row-wise application of the scalar function (the batched semantics)."""
    n = len(ag)
    out = np.empty(n, dtype=np.float32)
    for i in range(n):
        out[i] = compute_reward(ag[i], dg[i], ee_pos[i], ee_quat[i], width[i], task_index[i], **kw)
    return out


def threshold_adjacent(ag, dg, ee_pos, distance_threshold=0.05, tol=1e-6):
    """Rows whose d_place or d_reach lies within ``tol`` of the 0.05 thresholds: north_star
    exempts them from the bit-exact requirement (they are counted and reported instead)."""
    ag, dg, ee_pos = (np.asarray(x, dtype=np.float64) for x in (ag, dg, ee_pos))
    d_place = np.linalg.norm(ag - dg, axis=-1)
    d_reach = np.linalg.norm(ee_pos - ag, axis=-1)
    return (np.abs(d_place - distance_threshold) < tol) | (np.abs(d_reach - 0.05) < tol)
