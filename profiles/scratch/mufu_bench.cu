// Microbenchmark: MUFU.EX2 vs FFMA throughput per SM (B200).  nvcc -arch=sm_100a -O3 mufu_bench.cu -o mufu_bench
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, int iters, float seed) {
  float x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = seed + threadIdx.x * 1e-3f + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i])); }
      else if (MODE == 1) { x[i] = fmaf(x[i], 0.999f, 0.001f); }
      else if (MODE == 2) { asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(x[i])); }
      else if (MODE == 4) { unsigned u = __float_as_uint(x[i]); asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(u)); x[i] = __uint_as_float(u); }
      else if (MODE == 5) { unsigned u = __float_as_uint(x[i]); asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(u)); x[i] = __uint_as_float(u); }
      else { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i])); x[(i + 4) & 7] = fmaf(x[(i + 4) & 7], 0.999f, 0.001f); x[(i + 5) & 7] = fmaf(x[(i + 5) & 7], 0.999f, 0.001f); x[(i+6)&7] = fmaf(x[(i + 6) & 7], 0.999f, 0.001f);}
    }
  }
  float s = 0; for (int i = 0; i < 8; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char* name, int threads, int blocks_per_sm) {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int clk_khz; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  float* out; cudaMalloc(&out, sizeof(float) * sms * blocks_per_sm * threads);
  int iters = 20000;
  k<MODE><<<sms * blocks_per_sm, threads>>>(out, 100, 1.f);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a);
  k<MODE><<<sms * blocks_per_sm, threads>>>(out, iters, 1.f);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  double ops = (double)sms * blocks_per_sm * threads * iters * 8.0 * (MODE == 3 ? 1 : 1);
  printf("%-28s threads/SM %4d: %.1f Gop/s total, %.2f op/clk/SM at nominal %.0f MHz (%.3f ms)\n", name, threads * blocks_per_sm,
         ops / ms / 1e6, ops / (ms * 1e-3) / sms / (clk_khz * 1e3), clk_khz / 1e3, ms);
  cudaFree(out);
}
int main() {
  run<0>("MUFU.EX2", 256, 4); run<0>("MUFU.EX2", 128, 2); run<0>("MUFU.EX2 (1 warp/SMSP)", 128, 1);
  run<2>("MUFU.RCP", 256, 4);
  run<4>("EX2 bf16x2 (instr count; x2 elements)", 256, 4); run<4>("EX2 bf16x2, 1 warp/SMSP", 128, 1);
  run<5>("EX2 f16x2 (instr count; x2 elements)", 256, 4);
  run<1>("FFMA", 256, 4);
  run<3>("EX2 + 3 FFMA interleaved (EX2 count)", 256, 4); run<3>("EX2 + 3 FFMA, 1 warp/SMSP", 128, 1);
  return 0;
}
