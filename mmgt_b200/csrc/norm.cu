// GroupNorm (+SiLU, + virtual channel concat) and LayerNorm (+ positional encoding) -- HBM-bound kernels.
// Channels-last data: a frame is (T, C); group g owns channels [g*cpg, (g+1)*cpg) of every token.
#include "common.cuh"

namespace {

template <typename T> struct VecOf;
template <> struct VecOf<float> { static constexpr int N = 4; typedef float4 type; };
template <> struct VecOf<bf16> { static constexpr int N = 8; typedef uint4 type; };

template <typename T>
__device__ __forceinline__ void load_vec(const T* p, float (&v)[VecOf<T>::N]);
template <>
__device__ __forceinline__ void load_vec<float>(const float* p, float (&v)[4]) {
  float4 t = *reinterpret_cast<const float4*>(p);
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
template <>
__device__ __forceinline__ void load_vec<bf16>(const bf16* p, float (&v)[8]) {
  uint4 t = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 f = __bfloat1622float2(h[i]);
    v[2 * i] = f.x; v[2 * i + 1] = f.y;
  }
}
template <typename T>
__device__ __forceinline__ void store_vec(T* p, const float (&v)[VecOf<T>::N]);
template <>
__device__ __forceinline__ void store_vec<float>(float* p, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
template <>
__device__ __forceinline__ void store_vec<bf16>(bf16* p, const float (&v)[8]) {
  uint4 t;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = t;
}

constexpr int GN_THREADS = 256;
constexpr int GN_MAXNV = 3;  // channel vectors per thread when C/VEC > 256

// stats[(n*G + g)*2 + {0,1}] += {sum, sumsq} over the block's token chunk
template <typename T>
__global__ void __launch_bounds__(GN_THREADS)
gn_stats_kernel(const T* __restrict__ x1, const T* __restrict__ x2, double* __restrict__ stats, int T_tok, int C1,
                int C2, int groups, int tok_per_block) {
  constexpr int VEC = VecOf<T>::N;
  extern __shared__ float chan[];  // [2][C] per-channel partial sums of this block
  const int C = C1 + C2, Cv = C / VEC, C1v = C1 / VEC, cpg = C / groups;
  const int n = blockIdx.y;
  const int t0 = blockIdx.x * tok_per_block;
  const int t1 = min(T_tok, t0 + tok_per_block);
  for (int i = threadIdx.x; i < 2 * C; i += GN_THREADS) chan[i] = 0.f;
  __syncthreads();
  const int lanes = Cv < GN_THREADS ? Cv : GN_THREADS;
  const int rows_per_pass = GN_THREADS / lanes;
  const int lane = threadIdx.x % lanes, rip = threadIdx.x / lanes;
  float s[GN_MAXNV][VEC], ss[GN_MAXNV][VEC];
#pragma unroll
  for (int j = 0; j < GN_MAXNV; ++j)
#pragma unroll
    for (int e = 0; e < VEC; ++e) s[j][e] = ss[j][e] = 0.f;
  if (rip < rows_per_pass) {
    for (int t = t0 + rip; t < t1; t += rows_per_pass) {
      const size_t row = (size_t)n * T_tok + t;
#pragma unroll
      for (int j = 0; j < GN_MAXNV; ++j) {
        int cv = lane + j * lanes;
        if (cv < Cv) {
          float v[VEC];
          if (cv < C1v) load_vec<T>(x1 + row * C1 + (size_t)cv * VEC, v);
          else load_vec<T>(x2 + row * C2 + (size_t)(cv - C1v) * VEC, v);
#pragma unroll
          for (int e = 0; e < VEC; ++e) { s[j][e] += v[e]; ss[j][e] += v[e] * v[e]; }
        }
      }
    }
    // threads with the same lane (different token rows) share channels: <= rows_per_pass-way contention
#pragma unroll
    for (int j = 0; j < GN_MAXNV; ++j) {
      int cv = lane + j * lanes;
      if (cv < Cv) {
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
          atomicAdd(&chan[cv * VEC + e], s[j][e]);
          atomicAdd(&chan[C + cv * VEC + e], ss[j][e]);
        }
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < groups * 2; i += GN_THREADS) {
    const int g = i >> 1, which = i & 1;
    const float* src = chan + which * C + g * cpg;
    float a = 0.f;
    for (int c = 0; c < cpg; ++c) a += src[c];
    atomicAdd(&stats[(size_t)n * groups * 2 + i], (double)a);
  }
}

template <typename T>
__global__ void __launch_bounds__(GN_THREADS)
gn_apply_kernel(const T* __restrict__ x1, const T* __restrict__ x2, T* __restrict__ y, const float* __restrict__ gamma,
                const float* __restrict__ beta, const double* __restrict__ stats, int T_tok, int C1, int C2, int groups,
                float eps, int silu, int tok_per_block) {
  constexpr int VEC = VecOf<T>::N;
  extern __shared__ float sm[];  // scale[C], shift[C]
  const int C = C1 + C2, Cv = C / VEC, C1v = C1 / VEC, cpg = C / groups;
  float* scale = sm;
  float* shift = sm + C;
  const int n = blockIdx.y;
  const double cnt = (double)T_tok * cpg;
  for (int c = threadIdx.x; c < C; c += GN_THREADS) {
    int g = c / cpg;
    double mean = stats[((size_t)n * groups + g) * 2] / cnt;
    double var = stats[((size_t)n * groups + g) * 2 + 1] / cnt - mean * mean;
    if (var < 0) var = 0;
    float rstd = (float)(1.0 / sqrt(var + (double)eps));
    float sc = rstd * gamma[c];
    scale[c] = sc;
    shift[c] = beta[c] - (float)mean * sc;
  }
  __syncthreads();
  const int t0 = blockIdx.x * tok_per_block;
  const int t1 = min(T_tok, t0 + tok_per_block);
  const int lanes = Cv < GN_THREADS ? Cv : GN_THREADS;
  const int rows_per_pass = GN_THREADS / lanes;
  const int lane = threadIdx.x % lanes, rip = threadIdx.x / lanes;
  if (rip >= rows_per_pass) return;
  for (int t = t0 + rip; t < t1; t += rows_per_pass) {
    const size_t row = (size_t)n * T_tok + t;
    for (int cv = lane; cv < Cv; cv += lanes) {
      float v[VEC];
      if (cv < C1v) load_vec<T>(x1 + row * C1 + (size_t)cv * VEC, v);
      else load_vec<T>(x2 + row * C2 + (size_t)(cv - C1v) * VEC, v);
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        int c = cv * VEC + e;
        float o = v[e] * scale[c] + shift[c];
        v[e] = silu ? silu_f(o) : o;
      }
      store_vec<T>(y + row * C + (size_t)cv * VEC, v);
    }
  }
}

// one warp per row; the row lives in registers (two-pass mean / variance like torch)
template <typename T, int MAXIT>
__global__ void __launch_bounds__(256)
layernorm_kernel(const T* __restrict__ x, T* __restrict__ y, const float* __restrict__ gamma,
                 const float* __restrict__ beta, const float* __restrict__ pe, int64_t rows, int C, int T_tok, int F,
                 float eps) {
  constexpr int VEC = VecOf<T>::N;
  const int Cv = C / VEC;
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t row = warp; row < rows; row += nwarps) {
    float v[MAXIT][VEC];
    float sum = 0.f;
#pragma unroll
    for (int it = 0; it < MAXIT; ++it) {
      int cv = it * 32 + lane;
      if (cv < Cv) {
        load_vec<T>(x + row * C + (size_t)cv * VEC, v[it]);
#pragma unroll
        for (int e = 0; e < VEC; ++e) sum += v[it][e];
      }
    }
    const float mean = warp_sum(sum) / (float)C;
    float sq = 0.f;
#pragma unroll
    for (int it = 0; it < MAXIT; ++it) {
      int cv = it * 32 + lane;
      if (cv < Cv) {
#pragma unroll
        for (int e = 0; e < VEC; ++e) { float d = v[it][e] - mean; sq += d * d; }
      }
    }
    const float rstd = rsqrtf(warp_sum(sq) / (float)C + eps);
    const float* pe_row = pe ? pe + (size_t)((row / T_tok) % F) * C : nullptr;
#pragma unroll
    for (int it = 0; it < MAXIT; ++it) {
      int cv = it * 32 + lane;
      if (cv < Cv) {
        float o[VEC];
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
          int c = cv * VEC + e;
          o[e] = (v[it][e] - mean) * rstd * gamma[c] + beta[c];
          if (pe_row) o[e] += pe_row[c];
        }
        store_vec<T>(y + row * C + (size_t)cv * VEC, o);
      }
    }
  }
}

}  // namespace

extern "C" int mmgt_groupnorm(mmgt_ctx* ctx, const void* x1, const void* x2, void* y, const float* gamma,
                              const float* beta, double* stats_ws, int N, int T, int C1, int C2, int groups, float eps,
                              int silu, int dtype, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  MMGT_CHECK_ARG(x1 && y && gamma && beta && stats_ws && N > 0 && T > 0 && C1 > 0 && C2 >= 0 && groups > 0,
                 MMGT_E_INVALID, "groupnorm: bad args");
  MMGT_CHECK_ARG(C2 == 0 || x2, MMGT_E_INVALID, "groupnorm: x2 is NULL but C2 > 0");
  const int C = C1 + C2;
  MMGT_CHECK_ARG(C % groups == 0, MMGT_E_INVALID, "groupnorm: C=%d not divisible by groups=%d", C, groups);
  const int vec = dtype == MMGT_F32 ? 4 : 8;
  MMGT_CHECK_ARG(C1 % vec == 0 && C2 % vec == 0, MMGT_E_ALIGN, "groupnorm: channels must be multiples of %d", vec);
  MMGT_CHECK_ARG(C / vec <= GN_THREADS * GN_MAXNV, MMGT_E_UNSUPPORTED, "groupnorm: C=%d too large", C);
  MMGT_CHECK_ARG(aligned16(x1) && aligned16(y) && (!x2 || aligned16(x2)), MMGT_E_ALIGN, "groupnorm: 16B alignment");
  MMGT_CHECK_ARG(N <= 65535, MMGT_E_INVALID, "groupnorm: N too large");
  // ~8 blocks per SM worth of token chunks
  int chunks = (int)std::max<int64_t>(1, std::min<int64_t>(T, ((int64_t)ctx->num_sms * 8 + N - 1) / N));
  int tok_per_block = (T + chunks - 1) / chunks;
  int lanes = std::min(C / vec, GN_THREADS);
  int rpp = GN_THREADS / lanes;
  tok_per_block = std::max(tok_per_block, rpp);
  chunks = (T + tok_per_block - 1) / tok_per_block;
  MMGT_CUDA_OK(cudaMemsetAsync(stats_ws, 0, sizeof(double) * 2 * N * groups, st));
  dim3 grid(chunks, N);
  size_t smem_stats = sizeof(float) * 2 * C, smem_apply = sizeof(float) * 2 * C;
  MMGT_DISPATCH_DTYPE(dtype, T_, {
    gn_stats_kernel<T_><<<grid, GN_THREADS, smem_stats, st>>>((const T_*)x1, (const T_*)x2, stats_ws, T, C1, C2, groups,
                                                              tok_per_block);
    MMGT_LAUNCH_OK(ctx);
    gn_apply_kernel<T_><<<grid, GN_THREADS, smem_apply, st>>>((const T_*)x1, (const T_*)x2, (T_*)y, gamma, beta, stats_ws,
                                                              T, C1, C2, groups, eps, silu, tok_per_block);
    MMGT_LAUNCH_OK(ctx);
  });
  return 0;
}

extern "C" int mmgt_layernorm(mmgt_ctx* ctx, const void* x, void* y, const float* gamma, const float* beta,
                              const float* pe, int64_t rows, int C, int T, int F, float eps, int dtype, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  MMGT_CHECK_ARG(x && y && gamma && beta && rows > 0 && C > 0, MMGT_E_INVALID, "layernorm: bad args");
  MMGT_CHECK_ARG(!pe || (T > 0 && F > 0), MMGT_E_INVALID, "layernorm: pe needs T and F");
  const int vec = dtype == MMGT_F32 ? 4 : 8;
  MMGT_CHECK_ARG(C % vec == 0, MMGT_E_ALIGN, "layernorm: C must be a multiple of %d", vec);
  MMGT_CHECK_ARG(C <= 2048, MMGT_E_UNSUPPORTED, "layernorm: C=%d > 2048", C);
  MMGT_CHECK_ARG(aligned16(x) && aligned16(y), MMGT_E_ALIGN, "layernorm: 16B alignment");
  if (T <= 0) T = 1;
  if (F <= 0) F = 1;
  int64_t blocks64 = std::min<int64_t>((rows + 7) / 8, (int64_t)ctx->num_sms * 16);
  int blocks = (int)blocks64;
  if (dtype == MMGT_F32) {
    if (C <= 1280) layernorm_kernel<float, 10><<<blocks, 256, 0, st>>>((const float*)x, (float*)y, gamma, beta, pe, rows, C, T, F, eps);
    else layernorm_kernel<float, 16><<<blocks, 256, 0, st>>>((const float*)x, (float*)y, gamma, beta, pe, rows, C, T, F, eps);
  } else if (dtype == MMGT_BF16) {
    if (C <= 1280) layernorm_kernel<bf16, 5><<<blocks, 256, 0, st>>>((const bf16*)x, (bf16*)y, gamma, beta, pe, rows, C, T, F, eps);
    else layernorm_kernel<bf16, 8><<<blocks, 256, 0, st>>>((const bf16*)x, (bf16*)y, gamma, beta, pe, rows, C, T, F, eps);
  } else { mmgt_set_error("layernorm: bad dtype"); return MMGT_E_INVALID; }
  MMGT_LAUNCH_OK(ctx);
  return 0;
}
