"""Reward / success / goal-distance arithmetic of FrankaEnv on the GPU.

Mirrors the four methods of /root/reference/panda_mujoco_gym/envs/panda_env.py that sit on
the hot path, with the same names and argument meaning:

    compute_reward(achieved_goal, desired_goal, info) -> np.float32      panda_env.py:205-245
    _is_success(achieved_goal, desired_goal)          -> np.float32      panda_env.py:303-306
    goal_distance(a, b)                                                  panda_env.py:311-315
    get_ee_position / get_ee_orientation / get_fingers_width             panda_env.py:337-352

The reference's compute_reward is not a pure function of its arguments: it reads the EE
position, EE quaternion and finger width from the live simulator plus
``current_task_index``.  Here that hidden state is explicit:

* scalar call (reference shape: goals (3,)): the state comes from ``info`` keys ``ee_pos``,
  ``ee_quat``, ``fingers_width``, ``task_index`` when present, else from this object's
  attributes of the same meaning (set them like the simulator would).
* batched call (goals (N,3)): ``info`` is a dict of arrays ``ee_pos[N,3]``, ``ee_quat[N,4]``
  (wxyz), ``fingers_width[N]``, ``task_index[N]`` - or a sequence of N per-transition dicts,
  which is what SB3's ``HerReplayBuffer`` passes to ``env_method("compute_reward", ...)``.

Everything is evaluated by ``reward_kernel`` in libpnp_b200.so (FP64, reference operation order,
bit-exact).  NumPy inputs take the library's host pipeline; CUDA tensors stay on the device.
"""

from __future__ import annotations

from typing import Any, Mapping, Optional, Sequence

import numpy as np
import torch

from .. import engine

# panda_env.py:29-30 evaluated with gymnasium_robotics' euler2quat in float64
VERTICAL_QUAT = np.array([1.0, 0.0, -0.0, 0.0])
HORIZONTAL_QUAT = np.array([0.7071067811865476, -0.7071067811865475, 0.0, 0.0])

_STATE_KEYS = ("ee_pos", "ee_quat", "fingers_width", "task_index")


class FrankaRewardModel:
    VERTICAL_QUAT = VERTICAL_QUAT
    HORIZONTAL_QUAT = HORIZONTAL_QUAT

    def __init__(
        self,
        reward_type: str = "dense",
        distance_threshold: float = 0.05,
        task_sequence: Optional[Sequence[str]] = None,
        high_pick_z: float = 0.35,
        initial_object_height: float = 0.001,
        device: Optional[Any] = None,
    ):
        if reward_type not in ("dense", "sparse"):
            raise ValueError("reward_type must be 'dense' or 'sparse'")
        # same attribute names as FrankaEnv (panda_env.py:49-76, 139-141)
        self.reward_type = reward_type
        self.distance_threshold = distance_threshold
        self.task_sequence = list(task_sequence) if task_sequence is not None else ["cube1", "cube2", "cube3"]
        self.current_task_index = 0
        self.high_pick_z = high_pick_z
        self.initial_object_height = initial_object_height
        # hidden simulator state for scalar calls
        self.ee_pos = np.zeros(3)
        self.ee_quat = np.array([1.0, 0.0, 0.0, 0.0])
        self.fingers_width = 0.0
        self.device = torch.device(device) if device is not None else None
        self.last_counters: Optional[np.ndarray] = None
        self.dt = 0.002 * 25  # MujocoRobotEnv.dt = opt.timestep * n_substeps (shelf_pnp.py:19, shelf_pnp.xml:4)

    # ---- hidden-state getters (panda_env.py:337-352) -----------------------------------
    def get_ee_position(self):
        return self.ee_pos

    def get_ee_orientation(self):
        return self.ee_quat

    def get_fingers_width(self):
        return self.fingers_width

    # ---- helpers -------------------------------------------------------------------------
    def _params(self):
        key = (self.reward_type, len(self.task_sequence), self.initial_object_height, self.distance_threshold,
               self.high_pick_z)
        cached = getattr(self, "_params_cache", None)
        if cached is None or cached[0] != key:  # the attributes are public and may be reassigned between calls
            cached = self._params_cache = (key, engine.reward_params(*key))
        return cached[1]

    def _compute_reward_one(self, achieved_goal, desired_goal, info):
        """Scalar call (goals (3,), the shape FrankaEnv.step and test/reward_test.py use): one kernel launch through
        the library's mapped mailbox (pnp_reward_one_host_f64), FP64, bit-exact.  Returns np.float32."""
        st = getattr(self, "_one_stage", None)
        if st is None:
            st = self._one_stage = (np.empty(3), np.empty(3), np.empty(3), np.empty(4))
        ag, dg, ee, eq = st
        ag[:] = achieved_goal
        dg[:] = desired_goal
        if info:
            ee[:] = info["ee_pos"] if "ee_pos" in info else self.get_ee_position()
            eq[:] = info["ee_quat"] if "ee_quat" in info else self.get_ee_orientation()
            width = info["fingers_width"] if "fingers_width" in info else self.get_fingers_width()
            task = info["task_index"] if "task_index" in info else self.current_task_index
        else:
            ee[:] = self.get_ee_position()
            eq[:] = self.get_ee_orientation()
            width, task = self.get_fingers_width(), self.current_task_index
        params = self._params()
        if self.device is not None and torch.cuda.current_device() != self.device.index:
            with torch.cuda.device(self.device):
                r, succ, bits = engine.reward_one_host(ag, dg, ee, eq, width, task, params)
        else:
            r, succ, bits = engine.reward_one_host(ag, dg, ee, eq, width, task, params)
        self.last_counters = np.array([1, bits & 1, (bits >> 1) & 1, (bits >> 2) & 1], dtype=np.uint64)
        return r

    def _state_from_info(self, info, n: Optional[int]):
        """Collect ee_pos / ee_quat / fingers_width / task_index for n rows (None = scalar)."""
        if info is not None and not isinstance(info, Mapping):
            # sequence of per-transition dicts (SB3 HerReplayBuffer convention)
            infos = list(info)
            if n is None or len(infos) != n:
                raise ValueError("a sequence `info` must hold one dict per row")
            missing = [k for k in _STATE_KEYS if any(k not in d for d in infos)]
            if missing:
                raise ValueError(f"per-row info dicts lack {missing}: the reference reads these from the live sim")
            return {k: np.stack([np.asarray(d[k]) for d in infos]) for k in _STATE_KEYS}
        info = info or {}
        if n is None:
            return dict(
                ee_pos=info.get("ee_pos", self.get_ee_position()),
                ee_quat=info.get("ee_quat", self.get_ee_orientation()),
                fingers_width=info.get("fingers_width", self.get_fingers_width()),
                task_index=info.get("task_index", self.current_task_index),
            )
        missing = [k for k in _STATE_KEYS[:3] if k not in info]
        if missing:
            raise ValueError(
                f"batched compute_reward needs info[{missing}] (arrays with one row per transition): "
                "the reference reads them from the live simulator (panda_env.py:211-224)"
            )
        out = {k: info[k] for k in _STATE_KEYS[:3]}
        ti = info.get("task_index", None)
        if ti is None:
            ti = np.full((n,), self.current_task_index, dtype=np.int32)
        out["task_index"] = ti
        return out

    def _run(self, ag, dg, st, want_success):
        params = self._params()
        if isinstance(ag, torch.Tensor) and ag.is_cuda:
            dt = ag.dtype if ag.dtype in (torch.float32, torch.float64) else torch.float64
            dev = ag.device
            conv = lambda x, d=dt: torch.as_tensor(x, device=dev).to(d).contiguous()  # noqa: E731
            counters = torch.zeros(4, dtype=torch.int64, device=dev)
            rew, succ = engine.reward(
                conv(ag), conv(dg), conv(st["ee_pos"]), conv(st["ee_quat"]), conv(st["fingers_width"]),
                conv(st["task_index"], torch.int32), params, want_success=want_success, counters=counters,
            )
            self.last_counters = counters
            return rew, succ
        ag_np = np.asarray(ag)
        dt = np.float32 if ag_np.dtype == np.float32 else np.float64
        cast = lambda x, d=dt: np.ascontiguousarray(  # noqa: E731
            x.cpu().numpy() if isinstance(x, torch.Tensor) else x, dtype=d)
        if self.device is not None:
            ctx = torch.cuda.device(self.device)
        else:
            ctx = torch.cuda.device(torch.cuda.current_device()) if torch.cuda.is_available() else None
        if ctx is None:
            raise engine._lib.PnpLibraryError("no CUDA device: compute_reward has no CPU path")
        with ctx:
            rew, succ, counters = engine.reward_host(
                cast(ag), cast(dg), cast(st["ee_pos"]), cast(st["ee_quat"]), cast(st["fingers_width"]).reshape(-1),
                cast(st["task_index"], np.int32).reshape(-1), params, want_success=want_success,
            )
        self.last_counters = counters
        return rew, succ

    # ---- reference API -------------------------------------------------------------------
    def compute_reward(self, achieved_goal, desired_goal, info):
        """Scalar (goals (3,)) -> np.float32, or batched (goals (N,3)) -> float32[N]."""
        shape = tuple(achieved_goal.shape) if hasattr(achieved_goal, "shape") else np.shape(achieved_goal)
        if shape != (tuple(desired_goal.shape) if hasattr(desired_goal, "shape") else np.shape(desired_goal)):
            raise ValueError("achieved_goal and desired_goal must have the same shape")
        if len(shape) == 1:
            if shape != (3,):
                raise ValueError("goals must have shape (3,) or (N, 3)")
            if info is not None and not isinstance(info, Mapping):
                raise ValueError("a scalar compute_reward takes a dict `info` (or None)")
            if not torch.cuda.is_available():
                raise engine._lib.PnpLibraryError("no CUDA device: compute_reward has no CPU path")
            if isinstance(achieved_goal, torch.Tensor):
                achieved_goal = achieved_goal.detach().cpu().numpy()
            if isinstance(desired_goal, torch.Tensor):
                desired_goal = desired_goal.detach().cpu().numpy()
            return self._compute_reward_one(achieved_goal, desired_goal, info)
        if len(shape) != 2 or shape[1] != 3:
            raise ValueError("goals must have shape (3,) or (N, 3)")
        st = self._state_from_info(info, shape[0])
        rew, _ = self._run(achieved_goal, desired_goal, st, want_success=False)
        return rew

    def compute_reward_and_success(self, achieved_goal, desired_goal, info):
        """Batched: (reward float32[N], is_success float32[N]) from one pass over the rows."""
        st = self._state_from_info(info, int(achieved_goal.shape[0]))
        return self._run(achieved_goal, desired_goal, st, want_success=True)

    def _get_obs(self, state: Mapping[str, Any], tree=None, precision: str = "fp32"):
        """Batched FrankaEnv._get_obs (panda_env.py:279-301) from kinematic state.

        The reference reads the live simulator; here ``state`` carries what it reads, one row
        per env: ``q_arm[N,7]``, ``qvel_arm[N,7]``, ``fingers[N,2]`` (finger_joint1/2 qpos) and,
        for the current target cube, ``obj_pos[N,3]``, ``obj_quat[N,4]`` (wxyz), ``obj_vel[N,6]``
        (free-joint qvel: linear world, angular body-local), plus ``goal[N,3]`` or ``[3]``.
        Returns the reference's dict: ``observation`` (N,19), ``achieved_goal`` (N,3),
        ``desired_goal`` (N,3); CUDA tensors in -> CUDA tensors out, NumPy in -> NumPy out."""
        from ..tree import KinematicTree

        keys = ("q_arm", "qvel_arm", "fingers", "obj_pos", "obj_quat", "obj_vel", "goal")
        missing = [k for k in keys if k not in state]
        if missing:
            raise ValueError(f"_get_obs state lacks {missing}")
        on_gpu = isinstance(state["q_arm"], torch.Tensor) and state["q_arm"].is_cuda
        dt_t = torch.float32 if precision == "fp32" else torch.float64
        if not torch.cuda.is_available():
            raise engine._lib.PnpLibraryError("no CUDA device: _get_obs has no CPU path")
        dev = state["q_arm"].device if on_gpu else (self.device or torch.device("cuda", torch.cuda.current_device()))
        with torch.cuda.device(dev):
            engine.set_tree(tree or KinematicTree.from_mjcf())
            args = [torch.as_tensor(np.asarray(state[k]) if not isinstance(state[k], torch.Tensor) else state[k])
                    .to(device=dev, dtype=dt_t) for k in keys]
            rows = engine.get_obs(*args, dt=self.dt)
        out = {"observation": rows[:, :19], "achieved_goal": rows[:, 19:22], "desired_goal": rows[:, 22:25]}
        if not on_gpu:
            out = {k: v.double().cpu().numpy() for k, v in out.items()}
        return out

    def _is_success(self, achieved_goal, desired_goal):
        """np.float32(1.0 if ||ag - dg|| < distance_threshold else 0.0) (panda_env.py:303-306)."""
        d = self.goal_distance(achieved_goal, desired_goal)
        if isinstance(d, torch.Tensor):
            return (d < self.distance_threshold).to(torch.float32)
        if np.ndim(d) == 0:
            return np.float32(1.0 if float(d) < self.distance_threshold else 0.0)
        return (d < self.distance_threshold).astype(np.float32)

    def goal_distance(self, a, b):
        """np.linalg.norm(a - b, axis=-1) in float64 (panda_env.py:311-315), on the GPU."""
        if isinstance(a, torch.Tensor) and a.is_cuda:
            a2 = a.to(torch.float64).reshape(-1, 3)
            b2 = torch.as_tensor(b, device=a.device).to(torch.float64).reshape(-1, 3).expand_as(a2).contiguous()
            return engine.goal_distance(a2.contiguous(), b2).reshape(a.shape[:-1])
        a_np, b_np = np.array(a, dtype=np.float64), np.array(b, dtype=np.float64)
        a_b, b_b = np.broadcast_arrays(a_np, b_np)
        dev = self.device or torch.device("cuda", torch.cuda.current_device())
        ta = torch.as_tensor(np.ascontiguousarray(a_b).reshape(-1, 3), device=dev)
        tb = torch.as_tensor(np.ascontiguousarray(b_b).reshape(-1, 3), device=dev)
        d = engine.goal_distance(ta, tb).cpu().numpy().reshape(a_b.shape[:-1])
        return d if d.ndim else np.float64(d)
