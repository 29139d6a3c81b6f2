// Shared device-side definitions: constant-memory tree, scalar helpers, kinematics policies.
//
// Reference semantics implemented here (file:line under /root/reference):
//   mj_kinematics for the EE chain          panda_mujoco_gym/skills/ik_solver.py:58
//   mj_jacSite (hinge columns a x (p - c))  ik_solver.py:70-72
//   mju_mat2Quat                            panda_mujoco_gym/envs/panda_env.py:337-342
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/pnp_b200.h"
#include "pnp_vec.cuh"
#include "generated/spec_kinematics.cuh"

namespace pnp {

constexpr int NJ = PNP_NJOINT;

// Device copy of PnpTree in the compute precision.
template <typename T>
struct TreeDev {
  T link_pos[NJ * 3];
  T link_rot[NJ * 9];
  T ee_pos[3];
  T ee_rot[9];
  T lower[NJ];
  T upper[NJ];
  T qref[NJ];
};

__constant__ TreeDev<float> c_tree_f32;
__constant__ TreeDev<double> c_tree_f64;

template <typename T>
__device__ __forceinline__ const TreeDev<T>& ctree();
template <>
__device__ __forceinline__ const TreeDev<float>& ctree<float>() { return c_tree_f32; }
template <>
__device__ __forceinline__ const TreeDev<double>& ctree<double>() { return c_tree_f64; }

// ---------------------------------------------------------------------------------------------
// scalar helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void sincos_t(float x, float* s, float* c) { sincosf(x, s, c); }
__device__ __forceinline__ void sincos_t(double x, double* s, double* c) { sincos(x, s, c); }

// ---------------------------------------------------------------------------------------------
// Trig for the IK inner loops.  FP32: FIRST-order table look-up on a fine table (all FP32 IK kernels, the
// scalar-template ones through Trig<float>, the value-type ones through TrigV below - same arithmetic):
//   x = k*delta + r, delta = 2*pi/8192, k = rint(x/delta) by magic-number rounding, r by a 2-term Cody-Waite
//   reduction (|r| <= delta/2 = 3.83e-4);  sin x = sk + ck*r,  cos x = ck - sk*r   (truncation r^2/2 <= 7.4e-8).
// One table of sin(k*delta) with 8192 + 2048 entries (40 KB of shared memory): cos(k*delta) = sin((k + 2048)*delta)
// is the same table read 2048 entries further on, an immediate offset on the LDS.  8 instructions for the pair
// (sincosf: ~27 plus a Payne-Hanek branch), no F2I/I2F, no state; max abs error 1.3e-7, rms 2.9e-8 for |x| < 100 rad,
// 1.8e-7 up to 3.2e3 rad (tools/check_trig.py; the magic-number quadrant needs |x| * 1304 < 2^22: the FP32 IK
// kernels document |q| < 3.2e3 rad as their domain).  FP64 kernels call sincos().
// (The first session used a 1024-entry (sin, cos) table with a second-order correction: 9.1e-8 / 2.5e-8, but 12
// instructions per pair - on the packed path 10, four of them with three register-pair operands, which is what that
// pipe is short of.)
// ---------------------------------------------------------------------------------------------
constexpr int kTrigVN = 8192;
constexpr int kTrigVWords = kTrigVN + kTrigVN / 4;  // + a quarter turn for the cosine
__device__ __align__(16) float g_trigv_tab[kTrigVWords];  // sin(k * 2*pi/8192), filled by pnp_set_tree

template <typename T>
struct Trig;

template <>
struct Trig<float> {
  static constexpr bool kUsesTable = true;
  const float* tab;  // shared-memory copy of g_trigv_tab
  __device__ __forceinline__ void operator()(float x, float* s, float* c) const {
    const float t = fmaf(x, 1303.7972412109375f, 12582912.0f);
    const int ji = __float_as_int(t) & (kTrigVN - 1);
    const float k = t + (-12582912.0f);
    const float r = fmaf(k, -7.669904152862728e-4f, x);  // |k| <= 4900 inside the joint range: the rounding of the constant
                                                           // (2.1e-11) moves r by < 1.1e-7 rad - no second reduction term
    const float es = tab[ji], ec = tab[ji + kTrigVN / 4];
    *s = fmaf(ec, r, es);
    *c = fmaf(-es, r, ec);
  }
};

template <>
struct Trig<double> {
  static constexpr bool kUsesTable = false;
  const float* tab;  // unused
  __device__ __forceinline__ void operator()(double x, double* s, double* c) const { sincos(x, s, c); }
};

// Cooperative load of the trig table into shared memory (call by every thread of the block, __syncthreads() after).
// cp.async (LDGSTS): all of a thread's 16-byte copies are in flight at once and bypass the register file.  The plain
// LDG/STS loop ran in batches of four loads and took 5.5 us of a 32 us cfg2 launch (ncu: the STS waiting on its LDG
// held 17 % of the kernel's warp samples); a single-query solve paid the same 5 us.
__device__ __forceinline__ void load_trigv_table(float* s_tab) {
  const unsigned dst = (unsigned)__cvta_generic_to_shared(s_tab);
  for (int i = threadIdx.x; i < kTrigVWords / 4; i += blockDim.x)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 16u * (unsigned)i), "l"(g_trigv_tab + 4 * i) : "memory");
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

// one MUFU.RCP (callers pass pivots of J J^T + damping I in [damping, ~10] or distances > 1e-3: no
// range fix-up needed, which is what __fdividef(1, x) spends four more instructions on)
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float rcp_t(float x) { return rcp_approx(x); }
__device__ __forceinline__ double rcp_t(double x) { return 1.0 / x; }
__device__ __forceinline__ float sqrt_t(float x) { return sqrtf(x); }
__device__ __forceinline__ double sqrt_t(double x) { return sqrt(x); }
__device__ __forceinline__ float clamp_t(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
__device__ __forceinline__ double clamp_t(double x, double lo, double hi) { return fmin(fmax(x, lo), hi); }

// ---------------------------------------------------------------------------------------------
// Generic kinematics: any 7-hinge chain, tree read from __constant__ memory.
// J layout: J[r * 7 + j], rows 0-2 = jacp, rows 3-5 = jacr (fk_full only).
// ---------------------------------------------------------------------------------------------
struct GenericKin {
  static constexpr bool kSpecialized = false;

  template <typename T>
  static __device__ __forceinline__ T lower(int i) { return ctree<T>().lower[i]; }
  template <typename T>
  static __device__ __forceinline__ T upper(int i) { return ctree<T>().upper[i]; }
  template <typename T>
  static __device__ __forceinline__ T qref(int i) { return ctree<T>().qref[i]; }

  template <typename T, bool kWantJ, bool kWantRot>
  static __device__ __forceinline__ void chain(const T* __restrict__ s, const T* __restrict__ c,
                                               T* __restrict__ p_out, T* __restrict__ J,
                                               T* __restrict__ R_out) {
    const TreeDev<T>& t = ctree<T>();
    T R[9] = {T(1), T(0), T(0), T(0), T(1), T(0), T(0), T(0), T(1)};
    T p[3] = {T(0), T(0), T(0)};
    T anc[NJ][3], axs[NJ][3];
#pragma unroll
    for (int i = 0; i < NJ; ++i) {
      const T* lp = t.link_pos + 3 * i;
      const T* lr = t.link_rot + 9 * i;
#pragma unroll
      for (int r = 0; r < 3; ++r)
        p[r] = p[r] + (R[3 * r] * lp[0] + R[3 * r + 1] * lp[1] + R[3 * r + 2] * lp[2]);
      T N[9];
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int j = 0; j < 3; ++j)
          N[3 * r + j] = R[3 * r] * lr[j] + R[3 * r + 1] * lr[3 + j] + R[3 * r + 2] * lr[6 + j];
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        anc[i][r] = p[r];
        axs[i][r] = N[3 * r + 2];
        const T x = N[3 * r], y = N[3 * r + 1];
        R[3 * r] = c[i] * x + s[i] * y;       // frame * Rz(q): x' = c x + s y
        R[3 * r + 1] = c[i] * y - s[i] * x;   //                y' = c y - s x
        R[3 * r + 2] = N[3 * r + 2];
      }
    }
    T pe[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      pe[r] = p[r] + (R[3 * r] * t.ee_pos[0] + R[3 * r + 1] * t.ee_pos[1] + R[3 * r + 2] * t.ee_pos[2]);
      p_out[r] = pe[r];
    }
    if (kWantJ) {
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const T rx = pe[0] - anc[j][0], ry = pe[1] - anc[j][1], rz = pe[2] - anc[j][2];
        J[0 * 7 + j] = axs[j][1] * rz - axs[j][2] * ry;
        J[1 * 7 + j] = axs[j][2] * rx - axs[j][0] * rz;
        J[2 * 7 + j] = axs[j][0] * ry - axs[j][1] * rx;
        if (kWantRot) {
          J[3 * 7 + j] = axs[j][0];
          J[4 * 7 + j] = axs[j][1];
          J[5 * 7 + j] = axs[j][2];
        }
      }
    }
    if (kWantRot) {
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int j = 0; j < 3; ++j)
          R_out[3 * r + j] = R[3 * r] * t.ee_rot[j] + R[3 * r + 1] * t.ee_rot[3 + j] + R[3 * r + 2] * t.ee_rot[6 + j];
    }
  }

  template <typename T>
  static __device__ __forceinline__ void fk_pos(const T* s, const T* c, T* p) {
    chain<T, false, false>(s, c, p, nullptr, nullptr);
  }
  template <typename T>
  static __device__ __forceinline__ void fk_jacp(const T* s, const T* c, T* p, T* J) {
    chain<T, true, false>(s, c, p, J, nullptr);
  }
  template <typename T>
  static __device__ __forceinline__ void fk_full(const T* s, const T* c, T* p, T* J, T* R) {
    chain<T, true, true>(s, c, p, J, R);
  }
  // A (upper triangle a00 a01 a02 a11 a12 a22) = Jp Jp^T
  template <typename T>
  static __device__ __forceinline__ void jjt(const T* J, T* A) {
    const int rr[6] = {0, 0, 0, 1, 1, 2}, ss[6] = {0, 1, 2, 1, 2, 2};
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      T acc = T(0);
#pragma unroll
      for (int j = 0; j < NJ; ++j) acc = acc + J[rr[k] * 7 + j] * J[ss[k] * 7 + j];
      A[k] = acc;
    }
  }
  template <typename T>
  static __device__ __forceinline__ void jty(const T* J, const T* y, T* dq) {
#pragma unroll
    for (int j = 0; j < NJ; ++j) dq[j] = (J[j] * y[0] + J[7 + j] * y[1]) + J[14 + j] * y[2];
  }
};

// ---------------------------------------------------------------------------------------------
// Specialised kinematics: straight-line code generated for the packaged Panda tree
// (tools/gen_spec_kinematics.py).  Selected when the uploaded tree is bit-identical.
// ---------------------------------------------------------------------------------------------
struct SpecKin {
  static constexpr bool kSpecialized = true;

  template <typename T>
  static __device__ __forceinline__ T lower(int i) { return pnp_spec::spec_lower<T>(i); }
  template <typename T>
  static __device__ __forceinline__ T upper(int i) { return pnp_spec::spec_upper<T>(i); }
  template <typename T>
  static __device__ __forceinline__ T qref(int i) { return pnp_spec::spec_qref<T>(i); }

  template <typename T>
  static __device__ __forceinline__ void fk_pos(const T* s, const T* c, T* p) { pnp_spec::spec_fk_pos<T>(s, c, p); }
  template <typename T>
  static __device__ __forceinline__ void fk_jacp(const T* s, const T* c, T* p, T* J) {
    pnp_spec::spec_fk_jacp<T>(s, c, p, J);
  }
  template <typename T>
  static __device__ __forceinline__ void fk_full(const T* s, const T* c, T* p, T* J, T* R) {
#pragma unroll
    for (int k = 0; k < 42; ++k) J[k] = T(0);  // structural zeros are not written by the generator
    pnp_spec::spec_fk_full<T>(s, c, p, J, R);
  }
  template <typename T>
  static __device__ __forceinline__ void jjt(const T* J, T* A) { pnp_spec::spec_jjt<T>(J, A); }
  template <typename T>
  static __device__ __forceinline__ void jty(const T* J, const T* y, T* dq) { pnp_spec::spec_jty<T>(J, y, dq); }
};

// ---------------------------------------------------------------------------------------------
// mju_mat2Quat (engine_util_spatial.c), wxyz, normalised.  R row-major.
// ---------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ void mat2quat(const T* m, T* q) {
  const T half = T(0.5), quarter = T(0.25), one = T(1);
  if (m[0] + m[4] + m[8] > T(0)) {
    q[0] = half * sqrt_t(one + m[0] + m[4] + m[8]);
    const T k = quarter / q[0];
    q[1] = k * (m[7] - m[5]);
    q[2] = k * (m[2] - m[6]);
    q[3] = k * (m[3] - m[1]);
  } else if (m[0] > m[4] && m[0] > m[8]) {
    q[1] = half * sqrt_t(one + m[0] - m[4] - m[8]);
    const T k = quarter / q[1];
    q[0] = k * (m[7] - m[5]);
    q[2] = k * (m[1] + m[3]);
    q[3] = k * (m[2] + m[6]);
  } else if (m[4] > m[8]) {
    q[2] = half * sqrt_t(one - m[0] + m[4] - m[8]);
    const T k = quarter / q[2];
    q[0] = k * (m[2] - m[6]);
    q[1] = k * (m[1] + m[3]);
    q[3] = k * (m[5] + m[7]);
  } else {
    q[3] = half * sqrt_t(one - m[0] - m[4] + m[8]);
    const T k = quarter / q[3];
    q[0] = k * (m[3] - m[1]);
    q[1] = k * (m[2] + m[6]);
    q[2] = k * (m[5] + m[7]);
  }
  const T n = sqrt_t(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  const T inv = one / n;
  q[0] *= inv; q[1] *= inv; q[2] *= inv; q[3] *= inv;
}

// ---------------------------------------------------------------------------------------------
// One damped-least-squares evaluation (ik_solver.py:58-81) at joint angles q:
//   p  = FK(q)                      (mj_kinematics, :58-59)
//   n2 = |target - p|^2             (:60-61; caller takes the sqrt / compares)
//   qn = clip(q + clip(J^T (J J^T + damping I)^-1 e, +-step), lower, upper)   (:70-81)
// ---------------------------------------------------------------------------------------------
template <typename T>
struct IkConst {
  T pos_thresh, damping, step_limit;
  int max_iters;
};

template <typename T, typename Kin>
__device__ __forceinline__ void ik_eval_and_step(const T (&q)[NJ], const T (&tgt)[3], const IkConst<T>& k,
                                                 const Trig<T>& trig, T (&p)[3], T& n2, T (&qn)[NJ]) {
  T s[NJ], c[NJ];
#pragma unroll
  for (int i = 0; i < NJ; ++i) trig(q[i] - Kin::template qref<T>(i), &s[i], &c[i]);
  T J[21];
  Kin::template fk_jacp<T>(s, c, p, J);
  const T e0 = tgt[0] - p[0], e1 = tgt[1] - p[1], e2 = tgt[2] - p[2];
  n2 = (e0 * e0 + e1 * e1) + e2 * e2;
  T A[6];
  Kin::template jjt<T>(J, A);
  // (J J^T + damping I) y = e by LDL^T (the matrix is SPD for damping > 0)
  const T a00 = A[0] + k.damping, a11 = A[3] + k.damping, a22 = A[5] + k.damping;
  const T i0 = rcp_t(a00);
  const T l10 = A[1] * i0, l20 = A[2] * i0;
  const T d1 = a11 - l10 * A[1];
  const T u12 = A[4] - l10 * A[2];
  const T i1 = rcp_t(d1);
  const T l21 = u12 * i1;
  const T d2 = a22 - l20 * A[2] - l21 * u12;
  const T i2 = rcp_t(d2);
  const T z1 = e1 - l10 * e0;
  const T z2 = e2 - l20 * e0 - l21 * z1;
  T y[3];
  y[2] = z2 * i2;
  y[1] = z1 * i1 - l21 * y[2];
  y[0] = e0 * i0 - l10 * y[1] - l20 * y[2];
  T dq[NJ];
  Kin::template jty<T>(J, y, dq);
#pragma unroll
  for (int i = 0; i < NJ; ++i) {
    const T d = clamp_t(dq[i], -k.step_limit, k.step_limit);                              // :80
    qn[i] = clamp_t(q[i] + d, Kin::template lower<T>(i), Kin::template upper<T>(i));      // :81
  }
}

// ---------------------------------------------------------------------------------------------
// The same DLS evaluation in explicit-FMA form for the value types of pnp_vec.cuh: V = float (one
// query per lane) or V = F2 (two queries per lane on FFMA2/FMUL2/FADD2).  Every operation is the
// same correctly-rounded FP32 operation in both instantiations, so the packed kernel reproduces
// the scalar FP32 kernel bit for bit (tests/test_gpu_ik.py::test_pair_kernel_is_bit_identical).
// Specialised tree only (the generated "_v" kinematics).
// ---------------------------------------------------------------------------------------------
using pnp_spec::F2;
using pnp_spec::pnp_add;
using pnp_spec::pnp_fma;
using pnp_spec::pnp_mul;
using pnp_spec::pnp_neg;

__device__ __forceinline__ float v_sub(float a, float b) { return a - b; }
__device__ __forceinline__ F2 v_sub(F2 a, F2 b) { return pnp_spec::pnp_sub(a, b); }
__device__ __forceinline__ float v_rcp(float x) { return rcp_approx(x); }
__device__ __forceinline__ F2 v_rcp(F2 x) { return F2(rcp_approx(x.v.x), rcp_approx(x.v.y)); }
__device__ __forceinline__ float v_clamp_sym(float x, float lim) { return fminf(fmaxf(x, -lim), lim); }
__device__ __forceinline__ F2 v_clamp_sym(F2 x, F2 lim) {
  return F2(fminf(fmaxf(x.v.x, -lim.v.x), lim.v.x), fminf(fmaxf(x.v.y, -lim.v.y), lim.v.y));
}
__device__ __forceinline__ float v_clamp(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
__device__ __forceinline__ F2 v_clamp(F2 x, float lo, float hi) {
  return F2(fminf(fmaxf(x.v.x, lo), hi), fminf(fmaxf(x.v.y, lo), hi));
}

// Table trig of the value-type kernels: Trig<float>'s arithmetic for V = float and, packed, for V = F2 (the two
// slots of an F2 load straight into the halves of a register pair): 6 packed FP32 instructions per joint.
struct TrigV {
  const float* tab;  // shared memory copy of g_trigv_tab
  __device__ __forceinline__ void operator()(float x, float* s, float* c) const {
    const float t = fmaf(x, 1303.7972412109375f, 12582912.0f);
    const int ji = __float_as_int(t) & (kTrigVN - 1);
    const float k = t + (-12582912.0f);
    const float r = fmaf(k, -7.669904152862728e-4f, x);  // |k| <= 4900 inside the joint range: the rounding of the constant
                                                           // (2.1e-11) moves r by < 1.1e-7 rad - no second reduction term
    const float es = tab[ji], ec = tab[ji + kTrigVN / 4];
    *s = fmaf(ec, r, es);
    *c = fmaf(-es, r, ec);
  }
  __device__ __forceinline__ void operator()(F2 x, F2* s, F2* c) const {
    using P = F2;
    const P t = pnp_fma(x, P(1303.7972412109375f), P(12582912.0f));
    const int ja = __float_as_int(t.v.x) & (kTrigVN - 1), jb = __float_as_int(t.v.y) & (kTrigVN - 1);
    const P k = pnp_add(t, P(-12582912.0f));
    const P r = pnp_fma(k, P(-7.669904152862728e-4f), x);
    const P es(tab[ja], tab[jb]), ec(tab[ja + kTrigVN / 4], tab[jb + kTrigVN / 4]);
    *s = pnp_fma(ec, r, es);
    *c = pnp_fma(pnp_neg(es), r, ec);
  }
};

// Split in two so that a kernel can decide between the halves (from n2) whether a slot takes the step:
//   ik_eval_*  : p = FK(q), J = jacp, e = target - p, n2 = |e|^2          (ik_solver.py:58-61, 70-72)
//   ik_step_v  : qn = clip(q + clip(J^T (J J^T + damping I)^-1 e, +-slim), lower, upper)   (:78-81)
// slim is per slot: step_limit for a running query, 0 to freeze a finished one (qn == q).
//
// The frame of the evaluation.  J^T (J J^T + damping I)^-1 e and |e| do not depend on the frame J and e are written in, as
// long as they share it, and in the frame that JOINT 1 CARRIES (A_0 Rz(q_1) of the canonical chain) the first joint's
// rotation drops out of every product down the chain: FK + Jp is 85 operations instead of 107 (Panda), Jp has one more
// structural zero (3 operations off J J^T, 1 off J^T y), and what it costs is the rotation of the target into that frame,
// 4 operations per pass - 22 of 215 packed instructions per pass of the two-queries-per-lane kernel.  So:
//   tb          the target in the frame of joint 1's parent (spec_world_to_base_v: a constant rigid transform, once per query)
//   ik_eval_j1_v  everything in joint 1's frame: pr = FK there, J, e = Rz(-q_1) tb - pr; returns sin / cos of joint 1 too
//   p_world_v   pr back in the world frame (final_pos: only when a query is stored)
//   ik_eval_v   world-frame target in, world-frame p out (e and J are in joint 1's frame all the same)
// Every FP32 kernel on the specialised tree goes through ik_eval_j1_v, so they stay bit-identical to each other.
// e = Rz(-q_1) tb - pr and |e|^2: the error of a target (frame of joint 1's parent) at a position in joint 1's frame.
// Also what the fused waypoint / planner passes call for the NEXT solve's target at this pass's pr (same operations as
// the evaluation that solve's own first pass would do: bit-identical).
template <typename V>
__device__ __forceinline__ void target_err_j1_v(const V (&tb)[3], const V (&pr)[3], const V& s0, const V& c0, V (&e)[3], V& n2) {
  const V trx = pnp_fma(c0, tb[0], pnp_mul(s0, tb[1]));            // Rz(-q_1) tb
  const V try_ = pnp_fma(c0, tb[1], pnp_mul(pnp_neg(s0), tb[0]));
  e[0] = v_sub(trx, pr[0]); e[1] = v_sub(try_, pr[1]); e[2] = v_sub(tb[2], pr[2]);
  n2 = pnp_fma(e[2], e[2], pnp_fma(e[1], e[1], pnp_mul(e[0], e[0])));
}

template <typename V>
__device__ __forceinline__ void ik_eval_j1_v(const V (&q)[NJ], const V (&tb)[3], const TrigV& trig, V (&pr)[3],
                                             V (&e)[3], V& n2, V (&J)[21], V& s0, V& c0) {
  V s[NJ], c[NJ];
#pragma unroll
  for (int i = 0; i < NJ; ++i) {
    const float qr = pnp_spec::spec_qref<float>(i);
    trig(qr != 0.0f ? pnp_add(q[i], V(-qr)) : q[i], &s[i], &c[i]);
  }
  pnp_spec::spec_fk_jacp_j1_v<V>(s, c, pr, J);
  s0 = s[0]; c0 = c[0];
  target_err_j1_v<V>(tb, pr, s0, c0, e, n2);
}

// joint-1-frame position -> world (A_0.pos + A_0.rot Rz(q_1) pr)
template <typename V>
__device__ __forceinline__ void p_world_v(const V (&pr)[3], const V& s0, const V& c0, V (&pw)[3]) {
  V pb[3];
  pb[0] = pnp_fma(c0, pr[0], pnp_mul(pnp_neg(s0), pr[1]));
  pb[1] = pnp_fma(s0, pr[0], pnp_mul(c0, pr[1]));
  pb[2] = pr[2];
  pnp_spec::spec_base_to_world_v<V>(pb, pw);
}

template <typename V>
__device__ __forceinline__ void ik_eval_v(const V (&q)[NJ], const V (&tgt)[3], const TrigV& trig, V (&p)[3],
                                          V (&e)[3], V& n2, V (&J)[21], V& s0, V& c0) {
  V tb[3], pr[3];
  pnp_spec::spec_world_to_base_v<V>(tgt, tb);
  ik_eval_j1_v<V>(q, tb, trig, pr, e, n2, J, s0, c0);
  p_world_v<V>(pr, s0, c0, p);
}

template <typename V>
__device__ __forceinline__ void ik_step_v(V (&q)[NJ], const V (&J)[21], const V (&e)[3], float damping,
                                          const V& slim) {  // q is updated in place
  V A[6];
  pnp_spec::spec_jjt_damped_j1_v<V>(J, V(damping), A);  // J J^T + damping I (:76-77; damping is the start value of the diagonal sums)
  const V a00 = A[0], a11 = A[3], a22 = A[5];
  const V i0 = v_rcp(a00);
  const V l10 = pnp_mul(A[1], i0), l20 = pnp_mul(A[2], i0);
  const V d1 = pnp_fma(pnp_neg(l10), A[1], a11);
  const V u12 = pnp_fma(pnp_neg(l10), A[2], A[4]);
  const V i1 = v_rcp(d1);
  const V l21 = pnp_mul(u12, i1);
  const V d2 = pnp_fma(pnp_neg(l21), u12, pnp_fma(pnp_neg(l20), A[2], a22));
  const V i2 = v_rcp(d2);
  const V z1 = pnp_fma(pnp_neg(l10), e[0], e[1]);
  const V z2 = pnp_fma(pnp_neg(l21), z1, pnp_fma(pnp_neg(l20), e[0], e[2]));
  V y[3];
  y[2] = pnp_mul(z2, i2);
  y[1] = pnp_fma(pnp_neg(l21), y[2], pnp_mul(z1, i1));
  y[0] = pnp_fma(pnp_neg(l20), y[2], pnp_fma(pnp_neg(l10), y[1], pnp_mul(e[0], i0)));
  V dq[NJ];
  pnp_spec::spec_jty_j1_v<V>(J, y, dq);
#pragma unroll
  for (int i = 0; i < NJ; ++i) {
    // a joint whose Jacobian column is structurally zero (joint 7: the EE site lies on its axis) has
    // dq == 0 exactly: only the limit clip of :81 remains
    const bool moves = !pnp_spec::spec_jp_col_zero(i);
    const V qd = moves ? pnp_add(q[i], v_clamp_sym(dq[i], slim)) : q[i];                                 // :80
    q[i] = v_clamp(qd, pnp_spec::spec_lower<float>(i), pnp_spec::spec_upper<float>(i));                  // :81
  }
}

// FK only (final_pos of a solve, ik_solver.py:88)
template <typename T, typename Kin>
__device__ __forceinline__ void fk_position(const T (&q)[NJ], T (&p)[3]) {
  T s[NJ], c[NJ];
#pragma unroll
  for (int i = 0; i < NJ; ++i) sincos_t(q[i] - Kin::template qref<T>(i), &s[i], &c[i]);
  Kin::template fk_pos<T>(s, c, p);
}

}  // namespace pnp
