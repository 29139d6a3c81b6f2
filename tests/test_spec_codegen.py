"""The generated tree-specialised kinematics (host-side check of generated code, no GPU):
the committed header is up to date and, compiled as plain C++, agrees with the oracle."""
import ctypes
import os
import subprocess
import sys

import numpy as np

from conftest import ROOT
from oracle import c_oracle

GEN = os.path.join(ROOT, "tools", "gen_spec_kinematics.py")
HDR = os.path.join(ROOT, "mujoco_panda_pnp_b200", "csrc", "generated", "spec_kinematics.cuh")

_SHIM = r"""
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__ __restrict
#include "%s"
extern "C" void host_fk_full(const double* q, long n, double* p, double* J, double* R) {
  for (long i = 0; i < n; ++i) {
    double s[7], c[7];
    for (int k = 0; k < 7; ++k) { s[k] = __builtin_sin(q[7*i+k] - pnp_spec::spec_qref<double>(k)); c[k] = __builtin_cos(q[7*i+k] - pnp_spec::spec_qref<double>(k)); }
    for (int k = 0; k < 42; ++k) J[42*i+k] = 0.0;
    pnp_spec::spec_fk_full<double>(s, c, p + 3*i, J + 42*i, R + 9*i);
  }
}
extern "C" void host_fk_jacp(const double* q, long n, double* p, double* J, double* A, double* dq, const double* y) {
  for (long i = 0; i < n; ++i) {
    double s[7], c[7];
    for (int k = 0; k < 7; ++k) { s[k] = __builtin_sin(q[7*i+k]); c[k] = __builtin_cos(q[7*i+k]); }
    for (int k = 0; k < 21; ++k) J[21*i+k] = 0.0;
    pnp_spec::spec_fk_jacp<double>(s, c, p + 3*i, J + 21*i);
    pnp_spec::spec_jjt<double>(J + 21*i, A + 6*i);
    pnp_spec::spec_jty<double>(J + 21*i, y + 3*i, dq + 7*i);
    double p2[3];
    pnp_spec::spec_fk_pos<double>(s, c, p2);
    for (int k = 0; k < 3; ++k) if (p2[k] != p[3*i+k]) p[3*i+k] = 1e300;  // fk_pos must equal fk_jacp's p
  }
}
"""


def test_generated_header_is_up_to_date():
    assert subprocess.call([sys.executable, GEN, "--check"]) == 0, "run tools/gen_spec_kinematics.py"


def test_generated_code_matches_oracle_on_host(tmp_path, oracle_chain):
    src = tmp_path / "shim.cpp"
    src.write_text(_SHIM % HDR)
    so = tmp_path / "shim.so"
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", "-o", str(so), str(src)])
    lib = ctypes.CDLL(str(so))
    rng = np.random.default_rng(2)
    n = 2000
    q = rng.uniform(-3.0, 3.0, (n, 7))
    dp = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
    p, J, R = np.empty((n, 3)), np.empty((n, 6, 7)), np.empty((n, 3, 3))
    lib.host_fk_full(dp(q), ctypes.c_long(n), dp(p), dp(J), dp(R))
    pos, mat, jac = c_oracle.fk_jac(oracle_chain, q)
    np.testing.assert_allclose(p, pos, atol=2e-15 * 10)
    np.testing.assert_allclose(J, jac, atol=1e-14)
    np.testing.assert_allclose(R, mat, atol=1e-14)
    # position-only path + J J^T + J^T y helpers
    y = rng.normal(size=(n, 3))
    p2, J2, A, dq = np.empty((n, 3)), np.empty((n, 3, 7)), np.empty((n, 6)), np.empty((n, 7))
    lib.host_fk_jacp(dp(q), ctypes.c_long(n), dp(p2), dp(J2), dp(A), dp(dq), dp(y))
    np.testing.assert_allclose(p2, pos, atol=1e-14)
    np.testing.assert_allclose(J2, jac[:, :3], atol=1e-14)
    JJt = np.einsum("nij,nkj->nik", jac[:, :3], jac[:, :3])
    np.testing.assert_allclose(A, JJt[:, [0, 0, 0, 1, 1, 2], [0, 1, 2, 1, 2, 2]], atol=1e-13)
    np.testing.assert_allclose(dq, np.einsum("nij,ni->nj", jac[:, :3], y), atol=1e-13)
    # the structural zero the generator must have found: joint 7 does not move the EE site
    assert np.all(J2[:, :, 6] == 0.0)


_SHIM_J1 = r"""
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__ __restrict
namespace pnp_spec {
inline double pnp_fma(double a, double b, double c) { return a * b + c; }
inline double pnp_mul(double a, double b) { return a * b; }
inline double pnp_add(double a, double b) { return a + b; }
inline double pnp_neg(double a) { return -a; }
}
#include "%s"
// what ik_eval_j1_v / p_world_v / ik_step_v do with the generated pieces (pnp_common.cuh), in double
extern "C" void host_j1(const double* q, const double* tgt, const double* y, double lam, long n, double* p_world, double* J1,
                        double* e1, double* A, double* dq, double* sc0) {
  for (long i = 0; i < n; ++i) {
    double s[7], c[7];
    for (int k = 0; k < 7; ++k) { s[k] = __builtin_sin(q[7*i+k] - pnp_spec::spec_qref<double>(k)); c[k] = __builtin_cos(q[7*i+k] - pnp_spec::spec_qref<double>(k)); }
    double pr[3], J[21], tb[3], pb[3];
    for (int k = 0; k < 21; ++k) J[k] = 0.0;
    pnp_spec::spec_fk_jacp_j1_v<double>(s, c, pr, J);
    pnp_spec::spec_world_to_base_v<double>(tgt + 3*i, tb);
    e1[3*i]   = (c[0] * tb[0] + s[0] * tb[1]) - pr[0];
    e1[3*i+1] = (c[0] * tb[1] - s[0] * tb[0]) - pr[1];
    e1[3*i+2] = tb[2] - pr[2];
    pb[0] = c[0] * pr[0] - s[0] * pr[1]; pb[1] = s[0] * pr[0] + c[0] * pr[1]; pb[2] = pr[2];
    pnp_spec::spec_base_to_world_v<double>(pb, p_world + 3*i);
    for (int k = 0; k < 21; ++k) J1[21*i+k] = J[k];
    pnp_spec::spec_jjt_damped_j1_v<double>(J, lam, A + 6*i);
    pnp_spec::spec_jty_j1_v<double>(J, y + 3*i, dq + 7*i);
    sc0[2*i] = s[0]; sc0[2*i+1] = c[0];
  }
}
"""


def test_joint1_frame_kinematics_equal_the_world_frame_ones(tmp_path, oracle_chain):
    """spec_fk_jacp_j1_v (FK + Jp in the frame joint 1 carries, last site-moving joint folded into the offset) with the
    generated rigid transforms: rotated back, position and Jacobian are the oracle's; the error is the world error rotated;
    J J^T + lam I and J^T y are those of the rotated Jacobian with its structural zeros skipped."""
    src = tmp_path / "shim_j1.cpp"
    src.write_text(_SHIM_J1 % HDR)
    so = tmp_path / "shim_j1.so"
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", "-o", str(so), str(src)])
    lib = ctypes.CDLL(str(so))
    rng = np.random.default_rng(5)
    n = 2000
    q = rng.uniform(-3.0, 3.0, (n, 7))
    tgt = rng.uniform(-1.0, 2.0, (n, 3))
    y = rng.normal(size=(n, 3))
    lam = 0.01
    dp = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
    pw, J1, e1, A, dq, sc0 = (np.empty((n, 3)), np.empty((n, 3, 7)), np.empty((n, 3)), np.empty((n, 6)), np.empty((n, 7)),
                              np.empty((n, 2)))
    lib.host_j1(dp(q), dp(tgt), dp(y), ctypes.c_double(lam), ctypes.c_long(n), dp(pw), dp(J1), dp(e1), dp(A), dp(dq), dp(sc0))
    pos, _, jac = c_oracle.fk_jac(oracle_chain, q)
    np.testing.assert_allclose(pw, pos, atol=1e-14)
    # rotation world <- joint-1 frame: A_0.rot Rz(q_1); the packaged tree has A_0.rot = I (the generated transforms say so)
    s0, c0 = sc0[:, 0], sc0[:, 1]
    Rz = np.zeros((n, 3, 3))
    Rz[:, 0, 0], Rz[:, 0, 1], Rz[:, 1, 0], Rz[:, 1, 1], Rz[:, 2, 2] = c0, -s0, s0, c0, 1.0
    np.testing.assert_allclose(np.einsum("nij,njk->nik", Rz, J1), jac[:, :3], atol=1e-14)
    np.testing.assert_allclose(np.einsum("nij,nj->ni", Rz, e1), tgt - pos, atol=1e-14)
    JJt = np.einsum("nij,nkj->nik", J1, J1) + lam * np.eye(3)
    np.testing.assert_allclose(A, JJt[:, [0, 0, 0, 1, 1, 2], [0, 1, 2, 1, 2, 2]], atol=1e-13)
    np.testing.assert_allclose(dq, np.einsum("nij,ni->nj", J1, y), atol=1e-13)
    assert np.all(J1[:, :, 6] == 0.0) and np.all(J1[:, 1, 1] == 0.0) and np.all(J1[:, 2, 0] == 0.0)  # structural zeros
