"""ORACLE (test infrastructure): HER relabelling + VecNormalize as SB3 would apply them.

The reference only *promises* HER ("TQC(+HER)", /root/reference/scripts/train.py:4) and never wires
it, so there is no reference-side code for this step.  The semantics restated here
are those of the third-party packages the reference pins (stable-baselines3==2.2.1,
/root/reference/requirements.txt) - absent from /root/reference, not installable in this image:

* HerReplayBuffer._get_virtual_samples: new_goal = next_obs["achieved_goal"] of a later transition
  of the same episode; obs["desired_goal"] = next_obs["desired_goal"] = new_goal;
  reward = env.compute_reward(next_obs["achieved_goal"], new_goal, infos)
* VecNormalize.normalize_obs: clip((obs - mean) / sqrt(var + epsilon), -clip_obs, clip_obs), float32
  (clip_obs = 10, epsilon = 1e-8 in scripts/checkpoints/tqc_dense_vecnormalize_200000_steps.pkl)

The reward itself is the pinned restatement of FrankaEnv.compute_reward (reward_oracle.py) with the
hidden state taken from the stored next observation: ee_pos = observation[0:3], fingers_width =
observation[6] (panda_env.py:297), ee_quat / task_index from side arrays.
PARITY UNPINNED against real SB3 (nothing to run); reward bits are pinned through reward_oracle.
"""

from __future__ import annotations

import numpy as np

from . import reward_oracle


def relabel(obs, next_obs, future_idx, ee_quat, task_index, *, mean=None, var=None, epsilon=1e-8, clip_obs=10.0,
            **reward_kw):
    """obs/next_obs float32[N,25] rows (obs19 | ag3 | dg3).  Returns out_obs, out_next, reward, success."""
    obs = np.array(obs, dtype=np.float32)
    nxt = np.array(next_obs, dtype=np.float32)
    n = len(obs)
    goal = nxt[:, 22:25].copy()
    sel = np.asarray(future_idx) >= 0
    goal[sel] = np.asarray(next_obs, dtype=np.float32)[np.asarray(future_idx)[sel], 19:22]
    n64 = nxt.astype(np.float64)
    reward = reward_oracle.compute_reward_rows(n64[:, 19:22], goal.astype(np.float64), n64[:, 0:3],
                                               np.asarray(ee_quat, dtype=np.float64), n64[:, 6], task_index, **reward_kw)
    thr = reward_kw.get("distance_threshold", 0.05)
    success = np.array([reward_oracle.is_success(n64[i, 19:22], goal[i].astype(np.float64), thr) for i in range(n)],
                       dtype=np.float32)
    obs[:, 22:25] = goal
    nxt[:, 22:25] = goal
    if mean is not None:
        mean, var = np.asarray(mean, dtype=np.float64), np.asarray(var, dtype=np.float64)
        norm = lambda x: np.clip((x.astype(np.float64) - mean) / np.sqrt(var + epsilon), -clip_obs, clip_obs).astype(np.float32)  # noqa: E731
        obs, nxt = norm(obs), norm(nxt)
    return obs, nxt, reward, success
