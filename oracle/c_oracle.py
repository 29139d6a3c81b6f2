"""ORACLE (test infrastructure): ctypes front-end of oracle/c/pnp_oracle.c.

The C restatement is the multi-threaded FP64 "CPU path" that bench.py times beside the GPU
(cpu_baseline.kind = "port") and a second checker for the tests.
"""

from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "libpnp_oracle.so")
MAX_CHAIN = 16


class OracleChain(ctypes.Structure):
    _fields_ = [
        ("nbody", ctypes.c_int32),
        ("has_joint", ctypes.c_int32 * MAX_CHAIN),
        ("body_pos", ctypes.c_double * (MAX_CHAIN * 3)),
        ("body_quat", ctypes.c_double * (MAX_CHAIN * 4)),
        ("jnt_axis", ctypes.c_double * (MAX_CHAIN * 3)),
        ("jnt_pos", ctypes.c_double * (MAX_CHAIN * 3)),
        ("qpos0", ctypes.c_double * MAX_CHAIN),
        ("site_pos", ctypes.c_double * 3),
        ("site_quat", ctypes.c_double * 4),
        ("lower", ctypes.c_double * 7),
        ("upper", ctypes.c_double * 7),
    ]


class OracleIkParams(ctypes.Structure):
    _fields_ = [
        ("max_iters", ctypes.c_int32),
        ("pos_thresh", ctypes.c_double),
        ("damping", ctypes.c_double),
        ("step_limit", ctypes.c_double),
    ]


class OracleRewardParams(ctypes.Structure):
    _fields_ = [
        ("sparse", ctypes.c_int32),
        ("n_tasks", ctypes.c_int32),
        ("initial_object_height", ctypes.c_double),
        ("distance_threshold", ctypes.c_double),
        ("high_pick_z", ctypes.c_double),
    ]


class OracleMoveParams(ctypes.Structure):
    _fields_ = [
        ("pos_thresh", ctypes.c_double),
        ("max_traj_points", ctypes.c_int32),
        ("step_size", ctypes.c_double),
        ("max_outer", ctypes.c_int32),
        ("traj_cap", ctypes.c_int32),
    ]


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "c", "pnp_oracle.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.oracle_abi_version.restype = ctypes.c_int
    return _lib


def chain_from_model(model, site_name: str = "ee_center_site") -> OracleChain:
    """Raw MjModel-style body chain (no canonicalisation) world -> site body."""
    sid = model.site(site_name).id
    bodies = []
    b = int(model.site_bodyid[sid])
    while b != 0:
        bodies.append(b)
        b = int(model.body_parentid[b])
    bodies.reverse()
    c = OracleChain()
    c.nbody = len(bodies)
    assert len(bodies) <= MAX_CHAIN
    nj = 0
    for k, b in enumerate(bodies):
        c.body_pos[3 * k : 3 * k + 3] = list(map(float, model.body_pos[b]))
        c.body_quat[4 * k : 4 * k + 4] = list(map(float, model.body_quat[b]))
        jn, ja = int(model.body_jntnum[b]), int(model.body_jntadr[b])
        assert jn <= 1, "oracle chain supports at most one joint per body"
        c.has_joint[k] = jn
        if jn:
            assert int(model.jnt_type[ja]) == 3 and int(model.jnt_qposadr[ja]) == nj
            c.jnt_axis[3 * k : 3 * k + 3] = list(map(float, model.jnt_axis[ja]))
            c.jnt_pos[3 * k : 3 * k + 3] = list(map(float, model.jnt_pos[ja]))
            c.qpos0[k] = float(model.qpos0[nj])
            nj += 1
    assert nj == 7
    c.site_pos[:] = list(map(float, model.site_pos[sid]))
    c.site_quat[:] = list(map(float, model.site_quat[sid]))
    c.lower[:] = list(map(float, model.jnt_range[:7, 0]))
    c.upper[:] = list(map(float, model.jnt_range[:7, 1]))
    return c


def _dp(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def fk_jac(chain, q, nthreads=1):
    q = np.ascontiguousarray(q, dtype=np.float64).reshape(-1, 7)
    n = len(q)
    pos, mat, jac = np.empty((n, 3)), np.empty((n, 9)), np.empty((n, 6, 7))
    lib().oracle_fk_jac(ctypes.byref(chain), _dp(q), ctypes.c_int64(n), _dp(pos), _dp(mat), _dp(jac), int(nthreads))
    return pos, mat.reshape(n, 3, 3), jac


def ik_solve(chain, targets, q_init, max_iters=100, pos_thresh=1e-3, damping=1e-2, step_limit=0.1, nthreads=1):
    targets = np.ascontiguousarray(targets, dtype=np.float64).reshape(-1, 3)
    n = len(targets)
    q_init = np.ascontiguousarray(q_init, dtype=np.float64)
    stride = 0 if q_init.ndim == 1 else 7
    if stride:
        assert q_init.shape == (n, 7)
    p = OracleIkParams(int(max_iters), float(pos_thresh), float(damping), float(step_limit))
    q = np.empty((n, 7))
    fpos = np.empty((n, 3))
    err = np.empty(n)
    iters = np.empty(n, dtype=np.int32)
    flags = np.empty(n, dtype=np.uint8)
    lib().oracle_ik_solve(
        ctypes.byref(chain), ctypes.byref(p), _dp(targets), _dp(q_init), ctypes.c_int64(stride),
        ctypes.c_int64(n), _dp(q), _dp(fpos), _dp(err), _dp(iters), _dp(flags), int(nthreads),
    )
    return dict(q=q, final_pos=fpos, pos_error=err, iterations=iters,
                converged=(flags & 1).astype(bool), success=(flags & 2).astype(bool))


def reward(ag, dg, ee_pos, ee_quat, width, task_index, *, reward_type="dense", n_tasks=3,
           initial_object_height=0.001, distance_threshold=0.05, high_pick_z=0.35,
           want_success=True, nthreads=1):
    ag, dg, ee_pos = (np.ascontiguousarray(x, dtype=np.float64).reshape(-1, 3) for x in (ag, dg, ee_pos))
    ee_quat = np.ascontiguousarray(ee_quat, dtype=np.float64).reshape(-1, 4)
    width = np.ascontiguousarray(width, dtype=np.float64).reshape(-1)
    task_index = np.ascontiguousarray(task_index, dtype=np.int32).reshape(-1)
    n = len(ag)
    p = OracleRewardParams(int(reward_type == "sparse"), int(n_tasks), float(initial_object_height),
                           float(distance_threshold), float(high_pick_z))
    out = np.empty(n, dtype=np.float32)
    succ = np.empty(n, dtype=np.float32) if want_success else None
    lib().oracle_reward(
        ctypes.byref(p), _dp(ag), _dp(dg), _dp(ee_pos), _dp(ee_quat), _dp(width), _dp(task_index),
        ctypes.c_int64(n), _dp(out), _dp(succ) if want_success else None, int(nthreads),
    )
    return out, succ


def move_plan(chain, q_start, goal, pos_thresh=0.01, max_traj_points=200, step_size=0.01, max_outer=0, traj_cap=256,
              max_iters=100, ik_pos_thresh=1e-3, damping=1e-2, step_limit=0.1, nthreads=1):
    """MoveIKSkill.reset planner (skills/move.py:76-191) for n envs.  status bits: 1 = fallback
    strategies exhausted (reference `break`), 2 = max_outer reached, 4 = trajectory capacity hit."""
    q_start = np.ascontiguousarray(q_start, dtype=np.float64).reshape(-1, 7)
    goal = np.ascontiguousarray(goal, dtype=np.float64).reshape(-1, 3)
    n = len(q_start)
    assert len(goal) == n
    ikp = OracleIkParams(int(max_iters), float(ik_pos_thresh), float(damping), float(step_limit))
    mp = OracleMoveParams(float(pos_thresh), int(max_traj_points), float(step_size), int(max_outer), int(traj_cap))
    traj = np.zeros((n, traj_cap, 3))
    tlen = np.empty(n, dtype=np.int32)
    qf = np.empty((n, 7))
    solves = np.empty(n, dtype=np.int32)
    status = np.empty(n, dtype=np.int32)
    lib().oracle_move_plan(ctypes.byref(chain), ctypes.byref(ikp), ctypes.byref(mp), _dp(q_start), _dp(goal),
                           ctypes.c_int64(n), _dp(traj), _dp(tlen), _dp(qf), _dp(solves), _dp(status), int(nthreads))
    return dict(traj=traj, traj_len=tlen, q_final=qf, n_solves=solves, status=status)
