// Microbenchmark: does Blackwell's packed FP32 (FFMA2 / fma.rn.f32x2) relieve an issue-bound FP32 kernel?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp32x2_probe fp32x2_probe.cu && ./fp32x2_probe
// Prints, per variant, warp-instructions per clock per SM sub-partition and TFLOP/s.
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

constexpr int ITERS = 4096;

// A: 8 independent scalar FFMA chains
__global__ void k_ffma(float* out, float a, float b) {
  float x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 1e-3f + i;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = fmaf(x[i], a, b);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// B: 8 independent FFMA2 chains (16 floats)
__global__ void k_ffma2(float* out, float a, float b) {
  float2 x[8];
  const float2 aa = make_float2(a, a), bb = make_float2(b, b);
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f + i);
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = __ffma2_rn(x[i], aa, bb);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i].x + x[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// C: 8 FFMA + NALU FMNMX per iteration (scalar mix)
template <int NALU>
__global__ void k_ffma_alu(float* out, float a, float b, float lo) {
  float x[8], m[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { x[i] = threadIdx.x * 1e-3f + i; m[i] = x[i] * 0.5f; }
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = fmaf(x[i], a, b);
#pragma unroll
    for (int i = 0; i < NALU; ++i) m[i] = fmaxf(m[i], x[i]) + 0.0f * lo, m[i] = fminf(m[i], lo + i);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i] + m[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// D: 8 FFMA2 + NALU x (FMNMX pairs) per iteration
template <int NALU>
__global__ void k_ffma2_alu(float* out, float a, float b, float lo) {
  float2 x[8];
  float m[8];
  const float2 aa = make_float2(a, a), bb = make_float2(b, b);
#pragma unroll
  for (int i = 0; i < 8; ++i) { x[i] = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f + i); m[i] = x[i].x * 0.5f; }
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = __ffma2_rn(x[i], aa, bb);
#pragma unroll
    for (int i = 0; i < NALU; ++i) m[i] = fminf(fmaxf(m[i], x[i].x), lo + i);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i].x + x[i].y + m[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// E: dependent-chain latency (1 warp per SM sub-partition): scalar vs packed
__global__ void k_lat_ffma(float* out, float a, float b, long long* clk) {
  float x = threadIdx.x;
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) x = fmaf(x, a, b);
  }
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) clk[0] = t1 - t0;
}
__global__ void k_lat_ffma2(float* out, float a, float b, long long* clk) {
  float2 x = make_float2(threadIdx.x, threadIdx.x + 1);
  const float2 aa = make_float2(a, a), bb = make_float2(b, b);
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) x = __ffma2_rn(x, aa, bb);
  }
  long long t1 = clock64();
  out[threadIdx.x] = x.x + x.y;
  if (threadIdx.x == 0) clk[0] = t1 - t0;
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
  const int sms = prop.multiProcessorCount, threads = 256, blocks = sms * 8;
  float* out; CK(cudaMalloc(&out, sizeof(float) * blocks * threads));
  long long* clk; CK(cudaMalloc(&clk, 8));
  const double ghz = clk_khz * 1e-6;
  printf("device %s, %d SMs, max clock %.3f GHz\n", prop.name, sms, ghz);
  const double warps = (double)blocks * threads / 32.0;
  auto report = [&](const char* name, float ms, double fma_instr, double fma_width, double other_instr) {
    const double winstr = warps * ITERS * (fma_instr + other_instr);
    const double cycles = ms * 1e-3 * ghz * 1e9;
    const double ipc_smsp = winstr / cycles / (sms * 4.0);
    const double tflops = warps * ITERS * fma_instr * fma_width * 32 * 2 / (ms * 1e-3) * 1e-12;
    printf("%-34s %8.3f ms  %6.3f warp-instr/clk/SMSP (at max clock)  %7.2f TFLOP/s\n", name, ms, ipc_smsp, tflops);
  };
  float ms;
  ms = time_ms([&] { k_ffma<<<blocks, threads>>>(out, 0.999f, 0.001f); }); report("A  8xFFMA", ms, 8, 1, 0);
  ms = time_ms([&] { k_ffma2<<<blocks, threads>>>(out, 0.999f, 0.001f); }); report("B  8xFFMA2", ms, 8, 2, 0);
  ms = time_ms([&] { k_ffma_alu<4><<<blocks, threads>>>(out, 0.999f, 0.001f, 3.f); }); report("C4 8xFFMA + 4x(FMNMX,FMNMX)", ms, 8, 1, 8);
  ms = time_ms([&] { k_ffma2_alu<4><<<blocks, threads>>>(out, 0.999f, 0.001f, 3.f); }); report("D4 8xFFMA2 + 4x(FMNMX,FMNMX)", ms, 8, 2, 8);
  ms = time_ms([&] { k_ffma2_alu<8><<<blocks, threads>>>(out, 0.999f, 0.001f, 3.f); }); report("D8 8xFFMA2 + 8x(FMNMX,FMNMX)", ms, 8, 2, 16);
  long long h;
  k_lat_ffma<<<1, 32>>>(out, 0.999f, 0.001f, clk); CK(cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost));
  printf("latency FFMA  dependent: %.2f clk\n", (double)h / (ITERS * 8.0));
  k_lat_ffma2<<<1, 32>>>(out, 0.999f, 0.001f, clk); CK(cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost));
  printf("latency FFMA2 dependent: %.2f clk\n", (double)h / (ITERS * 8.0));
  CK(cudaDeviceSynchronize());
  return 0;
}
