// Value types the generated kinematics (generated/spec_kinematics.cuh, "_v" functions) and the DLS
// step are instantiated for:
//   float, double   - one query per lane
//   F2              - TWO queries per lane in one 64-bit register pair, arithmetic on Blackwell's
//                     packed FP32 instructions (PTX fma/mul/add.rn.f32x2 -> SASS FFMA2/FMUL2/FADD2).
// Why F2 exists: the FP32 pipe retires one warp-wide FFMA per clock per SM sub-partition, which is
// also the issue rate, so every non-FP32 instruction of a scalar kernel steals an FP32 slot
// (measured: issue slots 84 % busy, FMA pipe 54 %).  An FFMA2 occupies the pipe for two clocks but
// takes ONE issue slot (tools/microbench/fp32x2_probe.cu: 0.494 FFMA2/clk/SMSP = 73.6 TFLOP/s, same
// 4-clock dependent latency as FFMA), so the clamps, compares, table look-ups and control flow of
// two queries fit in the issue slots the packed math frees.  Operand negation and broadcast
// immediates are free modifiers on the packed instructions (checked in SASS).
#pragma once

#include <cuda_runtime.h>

namespace pnp_spec {

// kMode 0 (F2): every operation packed.  kMode 1 (F2H, "hybrid"): an FMA whose three operands are three
// DISTINCT register pairs (pnp_fma3, chosen by the generator / by hand where that is known at the source
// level) is issued as two scalar FFMAs instead of one FFMA2.  tools/microbench/fp32x2_operands.cu: an FFMA2
// reading three fresh register pairs holds the register file for 3 clocks (0.326 instr/clk = 48.5 TFLOP/s),
// two scalar FFMAs with three distinct registers each take 2.18 (0.916 instr/clk = 68.2 TFLOP/s); packed
// instructions with at most two fresh pairs (the rest immediate, reuse-cache or repeated) run at 2 clocks
// and stay packed.  Same correctly rounded FP32 operations either way: results are bit-identical.
template <int kMode>
struct F2T {
  float2 v;
  __device__ __forceinline__ F2T() {}
  __device__ __forceinline__ F2T(float a, float b) : v(make_float2(a, b)) {}
  __device__ __forceinline__ explicit F2T(float a) : v(make_float2(a, a)) {}
  __device__ __forceinline__ explicit F2T(double a) : v(make_float2((float)a, (float)a)) {}
  __device__ __forceinline__ explicit F2T(float2 a) : v(a) {}
  __device__ __forceinline__ float operator[](int k) const { return k == 0 ? v.x : v.y; }
  __device__ __forceinline__ void set(int k, float x) { if (k == 0) v.x = x; else v.y = x; }
};
using F2 = F2T<0>;
using F2H = F2T<1>;

__device__ __forceinline__ float pnp_fma(float a, float b, float c) { return fmaf(a, b, c); }
__device__ __forceinline__ float pnp_fma3(float a, float b, float c) { return fmaf(a, b, c); }
__device__ __forceinline__ float pnp_mul(float a, float b) { return a * b; }
__device__ __forceinline__ float pnp_add(float a, float b) { return a + b; }
__device__ __forceinline__ float pnp_neg(float a) { return -a; }

__device__ __forceinline__ double pnp_fma(double a, double b, double c) { return fma(a, b, c); }
__device__ __forceinline__ double pnp_fma3(double a, double b, double c) { return fma(a, b, c); }
__device__ __forceinline__ double pnp_mul(double a, double b) { return a * b; }
__device__ __forceinline__ double pnp_add(double a, double b) { return a + b; }
__device__ __forceinline__ double pnp_neg(double a) { return -a; }

template <int M>
__device__ __forceinline__ F2T<M> pnp_fma(F2T<M> a, F2T<M> b, F2T<M> c) { return F2T<M>(__ffma2_rn(a.v, b.v, c.v)); }
// three distinct register-pair operands
__device__ __forceinline__ F2T<0> pnp_fma3(F2T<0> a, F2T<0> b, F2T<0> c) { return F2T<0>(__ffma2_rn(a.v, b.v, c.v)); }
__device__ __forceinline__ F2T<1> pnp_fma3(F2T<1> a, F2T<1> b, F2T<1> c) {
  return F2T<1>(fmaf(a.v.x, b.v.x, c.v.x), fmaf(a.v.y, b.v.y, c.v.y));
}
template <int M>
__device__ __forceinline__ F2T<M> pnp_mul(F2T<M> a, F2T<M> b) { return F2T<M>(__fmul2_rn(a.v, b.v)); }
template <int M>
__device__ __forceinline__ F2T<M> pnp_add(F2T<M> a, F2T<M> b) { return F2T<M>(__fadd2_rn(a.v, b.v)); }
template <int M>
__device__ __forceinline__ F2T<M> pnp_neg(F2T<M> a) { return F2T<M>(-a.v.x, -a.v.y); }  // folds into the operand modifier
template <int M>
__device__ __forceinline__ F2T<M> pnp_sub(F2T<M> a, F2T<M> b) { return pnp_add(a, pnp_neg(b)); }

}  // namespace pnp_spec
