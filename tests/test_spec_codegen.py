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
