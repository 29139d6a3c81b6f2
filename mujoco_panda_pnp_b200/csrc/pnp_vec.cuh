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
// Measured and dropped (round 2): a "hybrid" pair type that issued every FMA with three distinct register-pair
// operands as two scalar FFMAs (the microbenchmark has an FFMA2 with three fresh pairs at 3 clocks and two such
// FFMAs at 2.2): 81 fewer FFMA2 and 162 more FFMA per pass, 3.17 ms against 3.04 ms for 2^24 queries - the issue
// slots the scalar instructions take cost more than the register-file cycles they save.
#pragma once

#include <cuda_runtime.h>

namespace pnp_spec {

struct F2 {
  float2 v;
  __device__ __forceinline__ F2() {}
  __device__ __forceinline__ F2(float a, float b) : v(make_float2(a, b)) {}
  __device__ __forceinline__ explicit F2(float a) : v(make_float2(a, a)) {}
  __device__ __forceinline__ explicit F2(double a) : v(make_float2((float)a, (float)a)) {}
  __device__ __forceinline__ explicit F2(float2 a) : v(a) {}
  __device__ __forceinline__ float operator[](int k) const { return k == 0 ? v.x : v.y; }
  __device__ __forceinline__ void set(int k, float x) { if (k == 0) v.x = x; else v.y = x; }
};

__device__ __forceinline__ float pnp_fma(float a, float b, float c) { return fmaf(a, b, c); }
__device__ __forceinline__ float pnp_mul(float a, float b) { return a * b; }
__device__ __forceinline__ float pnp_add(float a, float b) { return a + b; }
__device__ __forceinline__ float pnp_neg(float a) { return -a; }

__device__ __forceinline__ double pnp_fma(double a, double b, double c) { return fma(a, b, c); }
__device__ __forceinline__ double pnp_mul(double a, double b) { return a * b; }
__device__ __forceinline__ double pnp_add(double a, double b) { return a + b; }
__device__ __forceinline__ double pnp_neg(double a) { return -a; }

__device__ __forceinline__ F2 pnp_fma(F2 a, F2 b, F2 c) { return F2(__ffma2_rn(a.v, b.v, c.v)); }
__device__ __forceinline__ F2 pnp_mul(F2 a, F2 b) { return F2(__fmul2_rn(a.v, b.v)); }
__device__ __forceinline__ F2 pnp_add(F2 a, F2 b) { return F2(__fadd2_rn(a.v, b.v)); }
__device__ __forceinline__ F2 pnp_neg(F2 a) { return F2(-a.v.x, -a.v.y); }  // folds into the operand modifier
__device__ __forceinline__ F2 pnp_sub(F2 a, F2 b) { return pnp_add(a, pnp_neg(b)); }

}  // namespace pnp_spec
