// Microbenchmark (B200, sm_100a): do the FP64 pipe (DFMA) and the packed-FP32 pipe (FFMA2) run side by side?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_coissue fp64_coissue.cu && ./fp64_coissue
// Per variant: warp-instructions per clock per SM sub-partition, and the FFMA2 / DFMA parts of it.
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
constexpr int ITERS = 2048;

// NF independent FFMA2 chains (three distinct register pairs each if kThree) and ND independent DFMA chains per iteration
template <int NF, int ND, bool kThree>
__global__ void k_mix(float* out, float a, float b, double da, double db) {
  float2 x[NF > 0 ? NF : 1], y[NF > 0 ? NF : 1];
  double d[ND > 0 ? ND : 1], e[ND > 0 ? ND : 1];
  const float2 aa = make_float2(a, a), bb = make_float2(b, b);
#pragma unroll
  for (int i = 0; i < NF; ++i) { x[i] = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f + i); y[i] = make_float2(a + i, b - i); }
#pragma unroll
  for (int i = 0; i < ND; ++i) { d[i] = threadIdx.x * 1e-3 + i; e[i] = da + i; }
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < (NF > ND ? NF : ND); ++i) {
      if (i < NF) x[i] = kThree ? __ffma2_rn(x[i], y[i], y[(i + 1) % NF]) : __ffma2_rn(x[i], aa, bb);
      if (i < ND) d[i] = kThree ? fma(d[i], e[i], e[(i + 1) % ND]) : fma(d[i], da, db);
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < NF; ++i) s += x[i].x + x[i].y;
#pragma unroll
  for (int i = 0; i < ND; ++i) s += (float)d[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float time_ms(F launch) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  launch(); launch();
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 5; ++r) {
    cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  const int sms = prop.multiProcessorCount, threads = 128, blocks = sms * 4;  // 4 warps per scheduler, like the IK kernel
  float* out; CK(cudaMalloc(&out, sizeof(float) * blocks * threads));
  const double ghz = clk_khz * 1e-6;
  printf("device %s, %d SMs, max clock %.3f GHz, %d warps per scheduler\n", prop.name, sms, ghz, blocks * threads / 32 / sms / 4);
  const double warps = (double)blocks * threads / 32.0;
  auto report = [&](const char* name, float ms, int nf, int nd) {
    const double cycles = ms * 1e-3 * ghz * 1e9;
    const double f = warps * ITERS * nf / cycles / (sms * 4.0), d = warps * ITERS * nd / cycles / (sms * 4.0);
    printf("%-44s %8.3f ms  FFMA2 %5.3f + DFMA %5.3f = %5.3f warp-instr/clk/SMSP   FP32 %6.2f TFLOP/s  FP64 %6.2f TFLOP/s\n", name, ms, f, d, f + d,
           warps * ITERS * nf * 128.0 / (ms * 1e-3) * 1e-12, warps * ITERS * nd * 64.0 / (ms * 1e-3) * 1e-12);
  };
  float ms;
  ms = time_ms([&] { k_mix<8, 0, false><<<blocks, threads>>>(out, 0.999f, 0.001f, 0.999, 0.001); }); report("8 FFMA2 (1 pair + 2 shared)", ms, 8, 0);
  ms = time_ms([&] { k_mix<0, 8, false><<<blocks, threads>>>(out, 0.999f, 0.001f, 0.999, 0.001); }); report("8 DFMA (1 fresh + 2 shared)", ms, 0, 8);
  ms = time_ms([&] { k_mix<8, 8, false><<<blocks, threads>>>(out, 0.999f, 0.001f, 0.999, 0.001); }); report("8 FFMA2 + 8 DFMA interleaved", ms, 8, 8);
  ms = time_ms([&] { k_mix<8, 4, false><<<blocks, threads>>>(out, 0.999f, 0.001f, 0.999, 0.001); }); report("8 FFMA2 + 4 DFMA interleaved", ms, 8, 4);
  ms = time_ms([&] { k_mix<8, 0, true><<<blocks, threads>>>(out, 0.999f, 0.001f, 0.999, 0.001); }); report("8 FFMA2, three distinct pairs", ms, 8, 0);
  ms = time_ms([&] { k_mix<0, 8, true><<<blocks, threads>>>(out, 0.999f, 0.001f, 0.999, 0.001); }); report("8 DFMA, three distinct pairs", ms, 0, 8);
  ms = time_ms([&] { k_mix<8, 8, true><<<blocks, threads>>>(out, 0.999f, 0.001f, 0.999, 0.001); }); report("8 FFMA2 + 8 DFMA, three distinct pairs each", ms, 8, 8);
  ms = time_ms([&] { k_mix<8, 4, true><<<blocks, threads>>>(out, 0.999f, 0.001f, 0.999, 0.001); }); report("8 FFMA2 + 4 DFMA, three distinct pairs each", ms, 8, 4);
  CK(cudaDeviceSynchronize());
  return 0;
}
