// Microbenchmark 2: packed FP32 throughput as a function of how many DISTINCT register-pair operands
// an instruction reads (register-file port pressure).  See fp32x2_probe.cu for the basic rates.
#include <cstdio>
#include <cuda_runtime.h>
constexpr int ITERS = 4096;
constexpr int NC = 8;

template <int MODE>
__global__ void k(float* out, const float* in) {
  float2 x[NC], y[NC], z[NC];
  float xs[NC], ys[NC], zs[NC];
#pragma unroll
  for (int i = 0; i < NC; ++i) {
    x[i] = make_float2(in[i] + threadIdx.x, in[i + 8]);
    y[i] = make_float2(in[i + 16], in[i + 24] * 0.5f);
    z[i] = make_float2(in[i + 32], in[i + 40] * 0.25f);
    xs[i] = x[i].x; ys[i] = y[i].x; zs[i] = z[i].x;
  }
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < NC; ++i) {
      if (MODE == 0) x[i] = __ffma2_rn(x[i], y[0], z[0]);                 // 1 fresh pair + 2 reused pairs
      if (MODE == 1) x[i] = __ffma2_rn(y[i], z[i], x[i]);                 // 3 distinct pairs
      if (MODE == 2) x[i] = __ffma2_rn(y[i], z[(i + 3) % NC], x[i]);      // 3 distinct pairs, shuffled
      if (MODE == 3) x[i] = __fmul2_rn(x[i], y[i]);                       // 2 distinct pairs
      if (MODE == 4) xs[i] = fmaf(ys[i], zs[i], xs[i]);                   // scalar, 3 distinct regs
      if (MODE == 5) x[i] = __ffma2_rn(x[i], y[i], make_float2(0.316f, 0.316f));  // 2 pairs + immediate
      if (MODE == 6) x[i] = __ffma2_rn(x[i], make_float2(ys[i], ys[i]), z[i]);    // 2 pairs + scalar broadcast
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < NC; ++i) s += x[i].x + x[i].y + xs[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, float* out, const float* in, int sms, double ghz, int width) {
  const int threads = 256, blocks = sms * 4;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<blocks, threads>>>(out, in); k<MODE><<<blocks, threads>>>(out, in);
  float best = 1e30f;
  for (int r = 0; r < 5; ++r) {
    cudaEventRecord(e0); k<MODE><<<blocks, threads>>>(out, in); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  const double warps = (double)blocks * threads / 32.0;
  const double instr = warps * ITERS * NC;
  const double cyc = best * 1e-3 * ghz * 1e9;
  printf("%-44s %7.3f ms  %5.3f instr/clk/SMSP  %6.2f TFLOP/s\n", name, best, instr / cyc / (sms * 4.0),
         instr * 32 * 2 * width / (best * 1e-3) * 1e-12);
}

int main() {
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
  int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  const double ghz = clk_khz * 1e-6; const int sms = prop.multiProcessorCount;
  float *out, *in; cudaMalloc(&out, sizeof(float) * sms * 4 * 256); cudaMalloc(&in, 4096);
  float h[64]; for (int i = 0; i < 64; ++i) h[i] = 0.5f + 0.001f * i; cudaMemcpy(in, h, sizeof h, cudaMemcpyHostToDevice);
  printf("%s, %d SMs, %.3f GHz; 8 independent chains per thread, 8 warps per SM sub-partition\n", prop.name, sms, ghz);
  run<0>("FFMA2 x = x*y0+z0   (1 fresh pair, 2 reused)", out, in, sms, ghz, 2);
  run<1>("FFMA2 x = y_i*z_i+x (3 distinct pairs)", out, in, sms, ghz, 2);
  run<2>("FFMA2 x = y_i*z_j+x (3 distinct, shuffled)", out, in, sms, ghz, 2);
  run<3>("FMUL2 x = x*y_i     (2 distinct pairs)", out, in, sms, ghz, 1);
  run<4>("FFMA  x = y_i*z_i+x (scalar, 3 distinct)", out, in, sms, ghz, 1);
  run<5>("FFMA2 x = x*y_i+imm (2 pairs + immediate)", out, in, sms, ghz, 2);
  run<6>("FFMA2 x = x*bcast(y_i)+z_i (2 pairs + scalar)", out, in, sms, ghz, 2);
  cudaDeviceSynchronize();
  return 0;
}
