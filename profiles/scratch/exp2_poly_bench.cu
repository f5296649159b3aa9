// Round-2 experiment: how much softmax-exponential throughput does one SM gain when a share
// of the exp2 calls moves from MUFU.EX2 (16 / clk / SM) to the FMA pipe (Cody-Waite range reduction + degree-3 polynomial,
// relative error 7.7e-5 -- far below the bf16 rounding of P)?   nvcc -arch=sm_100a -O3 exp2_poly_bench.cu -o exp2_poly_bench
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>

__device__ __forceinline__ float ex2_mufu(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// x <= 0 (softmax argument in log2 units).  t = x + 1.5 * 2^23 rounds x to the nearest integer n in t's low mantissa
// bits; f = x - n in [-0.5, 0.5]; 2^x = p(f) * 2^n, the scaling done by adding n to the exponent field.
__device__ __forceinline__ float ex2_poly3(float x) {
  x = fmaxf(x, -126.f);
  const float t = x + 12582912.f;
  const float f = x - (t - 12582912.f);
  float p = fmaf(0.05508868396282196f, f, 0.24260404706001282f);
  p = fmaf(p, f, 0.6932762265205383f);
  p = fmaf(p, f, 0.9999289512634277f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

// EVERY-th exponential of each thread's 8-wide batch goes to the polynomial (EVERY = 0: all MUFU, 1: all polynomial)
template <int EVERY>
__global__ void k(float* out, int iters, float seed) {
  float x[8], acc = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = -(seed + threadIdx.x * 1e-3f + i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float a = fmaf(x[i], 0.999f, -1e-3f * it);      // the scale-and-subtract FFMA of the real softmax
      float e;
      if (EVERY != 0 && (EVERY == 1 || i % EVERY == 0)) e = ex2_poly3(a); else e = ex2_mufu(a);
      acc += e;
      x[i] = a + e * 1e-6f;
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

__global__ void accuracy(float* max_rel) {
  float worst = 0.f;
  for (int i = threadIdx.x; i < (1 << 22); i += blockDim.x) {
    const float x = -30.f * (float)i / (float)(1 << 22);
    const float ref = exp2f(x);
    worst = fmaxf(worst, fabsf(ex2_poly3(x) - ref) / ref);
  }
  atomicMax(reinterpret_cast<int*>(max_rel), __float_as_int(worst));
}

template <int EVERY> void run(const char* name, int threads, int blocks_per_sm) {
  int sms, clk_khz;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  float* out;
  cudaMalloc(&out, sizeof(float) * sms * blocks_per_sm * threads);
  const int iters = 20000;
  k<EVERY><<<sms * blocks_per_sm, threads>>>(out, 100, 1.f);
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a);
  k<EVERY><<<sms * blocks_per_sm, threads>>>(out, iters, 1.f);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  const double n = (double)sms * blocks_per_sm * threads * iters * 8.0;
  printf("%-34s threads/SM %4d: %.2f exp/clk/SM at nominal %.0f MHz (%.3f ms)\n", name, threads * blocks_per_sm,
         n / (ms * 1e-3) / sms / (clk_khz * 1e3), clk_khz / 1e3, ms);
  cudaFree(out);
}

int main() {
  float* d;
  cudaMalloc(&d, 4); cudaMemset(d, 0, 4);
  accuracy<<<1, 1024>>>(d);
  float h; cudaMemcpy(&h, d, 4, cudaMemcpyDeviceToHost);
  printf("ex2_poly3 max relative error on [-30, 0]: %.3e\n", h);
  run<0>("all MUFU.EX2", 256, 4);            run<0>("all MUFU.EX2, 2 warps/SMSP", 256, 1);
  run<4>("1 of 4 on the FMA pipe", 256, 4);  run<4>("1 of 4 on the FMA pipe, 2 warps/SMSP", 256, 1);
  run<2>("1 of 2 on the FMA pipe", 256, 4);  run<2>("1 of 2 on the FMA pipe, 2 warps/SMSP", 256, 1);
  run<1>("all on the FMA pipe", 256, 4);
  return 0;
}
