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

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// One kernel for statistics + normalise (+SiLU).  Work item = (frame n, token chunk); a block
//   1. accumulates per-channel sum / sum-of-squares of its chunk in registers -> shared -> per-group doubles in global,
//   2. bumps the frame's arrival counter and waits until all `chunks` items of the frame have arrived,
//   3. re-reads its chunk (now L2-resident: a frame is at most a few MB) and writes the normalised output.
// HBM traffic is one read + one write of the tensor instead of two reads + one write for the two-kernel form.
// Items are frame-major and the grid never exceeds the number of co-resident blocks, so every item a block waits
// for belongs to a block that is running (or has finished): no deadlock.
// NV = channel vectors per thread (1 when C / VEC <= 256), UNR = tokens in flight per thread.
template <typename T>
__device__ __forceinline__ void unpack_vec(const typename VecOf<T>::type& raw, float (&v)[VecOf<T>::N]);
template <>
__device__ __forceinline__ void unpack_vec<float>(const float4& t, float (&v)[4]) { v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
template <>
__device__ __forceinline__ void unpack_vec<bf16>(const uint4& t, float (&v)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 f = __bfloat1622float2(h[i]);
    v[2 * i] = f.x; v[2 * i + 1] = f.y;
  }
}

template <typename T, int NV>
__global__ void __launch_bounds__(GN_THREADS, 3)
gn_fused_kernel(const T* __restrict__ x1, const T* __restrict__ x2, T* __restrict__ y, const float* __restrict__ gamma,
                const float* __restrict__ beta, double* __restrict__ stats, unsigned* __restrict__ arrived, int N,
                int T_tok, int C1, int C2, int groups, float eps, int silu, int tok_per_block, int chunks) {
  constexpr int VEC = VecOf<T>::N;
  typedef typename VecOf<T>::type Raw;
  constexpr int UNR = 8 / NV < 1 ? 1 : 8 / NV;   // 16-byte loads in flight per thread = UNR * NV (kept packed)
  pdl_prologue();   // launched plainly (a memset node precedes it), but lets the NEXT kernel be scheduled early
  extern __shared__ float sm[];  // phase 1: [2][C] channel partial sums; phase 3: scale[C], shift[C]
  const int C = C1 + C2, Cv = C / VEC, C1v = C1 / VEC, cpg = C / groups;
  const int lanes = Cv < GN_THREADS ? Cv : GN_THREADS;
  const int rows_per_pass = GN_THREADS / lanes;
  const int lane = threadIdx.x % lanes, rip = threadIdx.x / lanes;
  const bool active = rip < rows_per_pass;
  const double cnt = (double)T_tok * cpg;
  const Raw zero_raw = {};

  auto src = [&](size_t row, int cv) -> const Raw* {
    return reinterpret_cast<const Raw*>(cv < C1v ? x1 + row * C1 + (size_t)cv * VEC : x2 + row * C2 + (size_t)(cv - C1v) * VEC);
  };

  for (int item = blockIdx.x; item < N * chunks; item += gridDim.x) {
    const int n = item / chunks;
    const int t0 = (item - n * chunks) * tok_per_block;
    const int t1 = min(T_tok, t0 + tok_per_block);
    // ---------------- phase 1: statistics
    for (int i = threadIdx.x; i < 2 * C; i += GN_THREADS) sm[i] = 0.f;
    __syncthreads();
    if (active) {
      float s[NV][VEC], ss[NV][VEC];
#pragma unroll
      for (int j = 0; j < NV; ++j)
#pragma unroll
        for (int e = 0; e < VEC; ++e) s[j][e] = ss[j][e] = 0.f;
      for (int t = t0 + rip; t < t1; t += UNR * rows_per_pass) {
        Raw raw[UNR][NV];
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
          const int tt = t + u * rows_per_pass;
          const size_t row = (size_t)n * T_tok + tt;
#pragma unroll
          for (int j = 0; j < NV; ++j) {
            const int cv = lane + j * lanes;
            raw[u][j] = (tt < t1 && cv < Cv) ? *src(row, cv) : zero_raw;
          }
        }
#pragma unroll
        for (int u = 0; u < UNR; ++u)
#pragma unroll
          for (int j = 0; j < NV; ++j) {
            float v[VEC];
            unpack_vec<T>(raw[u][j], v);
#pragma unroll
            for (int e = 0; e < VEC; ++e) { s[j][e] += v[e]; ss[j][e] += v[e] * v[e]; }
          }
      }
      // threads with the same lane (different token rows) share channels: <= rows_per_pass-way contention
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const int cv = lane + j * lanes;
        if (cv < Cv) {
#pragma unroll
          for (int e = 0; e < VEC; ++e) {
            atomicAdd(&sm[cv * VEC + e], s[j][e]);
            atomicAdd(&sm[C + cv * VEC + e], ss[j][e]);
          }
        }
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < groups * 2; i += GN_THREADS) {
      const int g = i >> 1, which = i & 1;
      const float* srcp = sm + which * C + g * cpg;
      float a = 0.f;
      for (int c = 0; c < cpg; ++c) a += srcp[c];
      atomicAdd(&stats[(size_t)n * groups * 2 + i], (double)a);
    }
    // ---------------- phase 2: frame barrier
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
      atomicAdd(&arrived[n], 1u);
      while (ld_acquire_u32(&arrived[n]) < (unsigned)chunks) __nanosleep(32);
    }
    __syncthreads();
    // ---------------- phase 3: normalise
    float* scale = sm;
    float* shift = sm + C;
    float* gstat = sm + 2 * C;      // (mean, rstd) per group: the double-precision part once per group, not per channel
    for (int g = threadIdx.x; g < groups; g += GN_THREADS) {
      const double mean = __ldcg(&stats[((size_t)n * groups + g) * 2]) / cnt;
      double var = __ldcg(&stats[((size_t)n * groups + g) * 2 + 1]) / cnt - mean * mean;
      if (var < 0) var = 0;
      gstat[2 * g] = (float)mean;
      gstat[2 * g + 1] = (float)(1.0 / sqrt(var + (double)eps));
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += GN_THREADS) {
      const int g = c / cpg;
      const float sc = gstat[2 * g + 1] * gamma[c];
      scale[c] = sc;
      shift[c] = beta[c] - gstat[2 * g] * sc;
    }
    __syncthreads();
    if (active) {
      for (int t = t0 + rip; t < t1; t += UNR * rows_per_pass) {
        Raw raw[UNR][NV];
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
          const int tt = t + u * rows_per_pass;
          const size_t row = (size_t)n * T_tok + tt;
#pragma unroll
          for (int j = 0; j < NV; ++j) {
            const int cv = lane + j * lanes;
            if (tt < t1 && cv < Cv) raw[u][j] = *src(row, cv);
          }
        }
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
          const int tt = t + u * rows_per_pass;
          const size_t row = (size_t)n * T_tok + tt;
#pragma unroll
          for (int j = 0; j < NV; ++j) {
            const int cv = lane + j * lanes;
            if (tt < t1 && cv < Cv) {
              float v[VEC];
              unpack_vec<T>(raw[u][j], v);
#pragma unroll
              for (int q = 0; q < VEC / 4; ++q) {
                const float4 sc = *reinterpret_cast<const float4*>(scale + cv * VEC + 4 * q);
                const float4 sh = *reinterpret_cast<const float4*>(shift + cv * VEC + 4 * q);
                float* o = &v[4 * q];
                o[0] = o[0] * sc.x + sh.x; o[1] = o[1] * sc.y + sh.y; o[2] = o[2] * sc.z + sh.z; o[3] = o[3] * sc.w + sh.w;
                if (silu) { o[0] = silu_f(o[0]); o[1] = silu_f(o[1]); o[2] = silu_f(o[2]); o[3] = silu_f(o[3]); }
              }
              store_vec<T>(y + row * C + (size_t)cv * VEC, v);
            }
          }
        }
      }
    }
    __syncthreads();   // sm is re-zeroed by the next item
  }
}

// ---- Two-kernel GroupNorm (default): statistics, then normalise (+SiLU).  No cross-CTA wait inside a kernel, so both
// run at full occupancy with every CTA streaming independently; the second read of the tensor is served by the 126 MB
// L2 (a level-0 activation is 63 MB, the deeper levels 4-31 MB).  Measured against the single fused kernel above
// (spin barrier per frame, 35 % occupancy, 26 % of HBM peak at C = 320): profiles/r2_*.
__device__ __forceinline__ float silu_fast(float v) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(v * -1.4426950408889634f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.f));
  return v * r;
}

template <typename T, int NV>
__global__ void __launch_bounds__(GN_THREADS, 3)
gn_stats_kernel(const T* __restrict__ x1, const T* __restrict__ x2, double* __restrict__ stats, int T_tok, int C1, int C2,
                int groups, int tok_per_block) {
  constexpr int VEC = VecOf<T>::N;
  typedef typename VecOf<T>::type Raw;
  constexpr int UNR = 8 / NV < 1 ? 1 : 8 / NV;
  pdl_prologue();
  extern __shared__ float sm[];   // [2][rows_per_pass][C] per-thread partial sums
  const int C = C1 + C2, Cv = C / VEC, C1v = C1 / VEC, cpg = C / groups;
  const int lanes = Cv < GN_THREADS ? Cv : GN_THREADS;
  const int rows_per_pass = GN_THREADS / lanes;
  const int lane = threadIdx.x % lanes, rip = threadIdx.x / lanes;
  const bool active = rip < rows_per_pass;
  const int n = blockIdx.y;
  const int t0 = blockIdx.x * tok_per_block;
  const int t1 = min(T_tok, t0 + tok_per_block);
  const Raw zero_raw = {};
  auto src = [&](size_t row, int cv) -> const Raw* {
    return reinterpret_cast<const Raw*>(cv < C1v ? x1 + row * C1 + (size_t)cv * VEC : x2 + row * C2 + (size_t)(cv - C1v) * VEC);
  };
  if (active) {
    float s[NV][VEC], ss[NV][VEC];
#pragma unroll
    for (int j = 0; j < NV; ++j)
#pragma unroll
      for (int e = 0; e < VEC; ++e) s[j][e] = ss[j][e] = 0.f;
    for (int t = t0 + rip; t < t1; t += UNR * rows_per_pass) {
      Raw raw[UNR][NV];
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const int tt = t + u * rows_per_pass;
        const size_t row = (size_t)n * T_tok + tt;
#pragma unroll
        for (int j = 0; j < NV; ++j) {
          const int cv = lane + j * lanes;
          raw[u][j] = (tt < t1 && cv < Cv) ? *src(row, cv) : zero_raw;
        }
      }
#pragma unroll
      for (int u = 0; u < UNR; ++u)
#pragma unroll
        for (int j = 0; j < NV; ++j) {
          float v[VEC];
          unpack_vec<T>(raw[u][j], v);
#pragma unroll
          for (int e = 0; e < VEC; ++e) { s[j][e] += v[e]; ss[j][e] = fmaf(v[e], v[e], ss[j][e]); }
        }
    }
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int cv = lane + j * lanes;
      if (cv < Cv) {
        float* d0 = sm + (size_t)rip * C + cv * VEC;
        float* d1 = sm + (size_t)(rows_per_pass + rip) * C + cv * VEC;
#pragma unroll
        for (int e = 0; e < VEC; ++e) { d0[e] = s[j][e]; d1[e] = ss[j][e]; }
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < groups * 2; i += GN_THREADS) {
    const int g = i >> 1, which = i & 1;
    float a = 0.f;
    for (int r = 0; r < rows_per_pass; ++r) {
      const float* srcp = sm + (size_t)(which * rows_per_pass + r) * C + g * cpg;
      for (int c = 0; c < cpg; ++c) a += srcp[c];
    }
    atomicAdd(&stats[((size_t)n * groups + g) * 2 + which], (double)a);
  }
}

template <typename T, int NV>
__global__ void __launch_bounds__(GN_THREADS, 3)
gn_apply_kernel(const T* __restrict__ x1, const T* __restrict__ x2, T* __restrict__ y, const float* __restrict__ gamma,
                const float* __restrict__ beta, const double* __restrict__ stats, int T_tok, int C1, int C2, int groups,
                float eps, int silu, int tok_per_block) {
  constexpr int VEC = VecOf<T>::N;
  typedef typename VecOf<T>::type Raw;
  constexpr int UNR = 8 / NV < 1 ? 1 : 8 / NV;
  pdl_prologue();
  extern __shared__ float sm[];   // scale[C], shift[C]
  const int C = C1 + C2, Cv = C / VEC, C1v = C1 / VEC, cpg = C / groups;
  const int lanes = Cv < GN_THREADS ? Cv : GN_THREADS;
  const int rows_per_pass = GN_THREADS / lanes;
  const int lane = threadIdx.x % lanes, rip = threadIdx.x / lanes;
  const bool active = rip < rows_per_pass;
  const int n = blockIdx.y;
  const int t0 = blockIdx.x * tok_per_block;
  const int t1 = min(T_tok, t0 + tok_per_block);
  const double cnt = (double)T_tok * cpg;
  float* scale = sm;
  float* shift = sm + C;
  float* gstat = sm + 2 * C;        // (mean, rstd) per group: the double-precision part once per group, not per channel
  for (int g = threadIdx.x; g < groups; g += GN_THREADS) {
    const double mean = stats[((size_t)n * groups + g) * 2] / cnt;
    double var = stats[((size_t)n * groups + g) * 2 + 1] / cnt - mean * mean;
    if (var < 0) var = 0;
    gstat[2 * g] = (float)mean;
    gstat[2 * g + 1] = (float)(1.0 / sqrt(var + (double)eps));
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += GN_THREADS) {
    const int g = c / cpg;
    const float sc = gstat[2 * g + 1] * gamma[c];
    scale[c] = sc;
    shift[c] = beta[c] - gstat[2 * g] * sc;
  }
  __syncthreads();
  auto src = [&](size_t row, int cv) -> const Raw* {
    return reinterpret_cast<const Raw*>(cv < C1v ? x1 + row * C1 + (size_t)cv * VEC : x2 + row * C2 + (size_t)(cv - C1v) * VEC);
  };
  if (!active) return;
  // this thread's channels never change: keep their scale / shift in registers
  float sc[NV][VEC], sh[NV][VEC];
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int cv = lane + j * lanes;
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      sc[j][e] = cv < Cv ? scale[cv * VEC + e] : 0.f;
      sh[j][e] = cv < Cv ? shift[cv * VEC + e] : 0.f;
    }
  }
  for (int t = t0 + rip; t < t1; t += UNR * rows_per_pass) {
    Raw raw[UNR][NV];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int tt = t + u * rows_per_pass;
      const size_t row = (size_t)n * T_tok + tt;
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const int cv = lane + j * lanes;
        if (tt < t1 && cv < Cv) raw[u][j] = *src(row, cv);
      }
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int tt = t + u * rows_per_pass;
      const size_t row = (size_t)n * T_tok + tt;
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const int cv = lane + j * lanes;
        if (tt < t1 && cv < Cv) {
          float v[VEC];
          unpack_vec<T>(raw[u][j], v);
#pragma unroll
          for (int e = 0; e < VEC; ++e) {
            v[e] = fmaf(v[e], sc[j][e], sh[j][e]);
            if (silu) v[e] = silu_fast(v[e]);
          }
          store_vec<T>(y + row * C + (size_t)cv * VEC, v);
        }
      }
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


// Full-width fast path: a row is split over G = 8 / 16 / 32 lanes with VPL 16-byte vectors per lane, so one warp
// normalises 32 / G rows at a time and every lane has VPL independent loads in flight (the one-warp-per-row
// kernel below issues 1-2 loads per lane between two dependent shuffle reductions: latency-bound at C = 320).
// gamma / beta sit in shared memory and are read as float4.
// The row a lane group works on next is fetched (as raw 16-byte vectors) before the arithmetic and the stores of the current
// one, so with an exact-wave persistent grid (flag 17) every resident warp always has VPL loads in flight.
template <typename T, int VPL, int G, bool PREFETCH>
__global__ void __launch_bounds__(256)
layernorm_grp_kernel(const T* __restrict__ x, T* __restrict__ y, const float* __restrict__ gamma,
                     const float* __restrict__ beta, const float* __restrict__ pe, int64_t rows, int C, int T_tok, int F,
                     float eps) {
  constexpr int VEC = VecOf<T>::N;
  typedef typename VecOf<T>::type Raw;
  extern __shared__ float gb[];   // gamma[C], beta[C]
  pdl_prologue();
  const int gl = threadIdx.x % G;
  const int64_t grp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
  const int64_t ngrp = (int64_t)gridDim.x * blockDim.x / G;
  Raw nxt[VPL];
  if (PREFETCH && grp < rows) {
#pragma unroll
    for (int j = 0; j < VPL; ++j) nxt[j] = *reinterpret_cast<const Raw*>(x + grp * C + (size_t)(gl + j * G) * VEC);
  }
  for (int i = threadIdx.x; i < C; i += blockDim.x) { gb[i] = gamma[i]; gb[C + i] = beta[i]; }
  __syncthreads();
  const float inv_c = 1.f / (float)C;
  for (int64_t row = grp; row < rows; row += ngrp) {
    float v[VPL][VEC];
    if constexpr (PREFETCH) {
#pragma unroll
      for (int j = 0; j < VPL; ++j) unpack_vec<T>(nxt[j], v[j]);
      if (row + ngrp < rows) {
        const T* xn = x + (row + ngrp) * C;
#pragma unroll
        for (int j = 0; j < VPL; ++j) nxt[j] = *reinterpret_cast<const Raw*>(xn + (size_t)(gl + j * G) * VEC);
      }
    } else {
      const T* xr = x + row * C;
#pragma unroll
      for (int j = 0; j < VPL; ++j) load_vec<T>(xr + (size_t)(gl + j * G) * VEC, v[j]);
    }
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < VPL; ++j)
#pragma unroll
      for (int e = 0; e < VEC; ++e) sum += v[j][e];
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum * inv_c;
    float sq = 0.f;
#pragma unroll
    for (int j = 0; j < VPL; ++j)
#pragma unroll
      for (int e = 0; e < VEC; ++e) { const float d = v[j][e] - mean; sq += d * d; }
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    const float rstd = rsqrtf(sq * inv_c + eps);
    const float* pe_row = pe ? pe + (size_t)((row / T_tok) % F) * C : nullptr;
    T* yr = y + row * C;
#pragma unroll
    for (int j = 0; j < VPL; ++j) {
      const int c0 = (gl + j * G) * VEC;
      float o[VEC];
#pragma unroll
      for (int q = 0; q < VEC / 4; ++q) {
        const float4 g4 = *reinterpret_cast<const float4*>(gb + c0 + 4 * q);
        const float4 b4 = *reinterpret_cast<const float4*>(gb + C + c0 + 4 * q);
        o[4 * q + 0] = (v[j][4 * q + 0] - mean) * rstd * g4.x + b4.x;
        o[4 * q + 1] = (v[j][4 * q + 1] - mean) * rstd * g4.y + b4.y;
        o[4 * q + 2] = (v[j][4 * q + 2] - mean) * rstd * g4.z + b4.z;
        o[4 * q + 3] = (v[j][4 * q + 3] - mean) * rstd * g4.w + b4.w;
        if (pe_row) {
          const float4 p4 = __ldg(reinterpret_cast<const float4*>(pe_row + c0 + 4 * q));
          o[4 * q + 0] += p4.x; o[4 * q + 1] += p4.y; o[4 * q + 2] += p4.z; o[4 * q + 3] += p4.w;
        }
      }
      store_vec<T>(yr + c0, o);
    }
  }
}

template <typename T, int VPL, int G>
void launch_ln_grp(mmgt_ctx* ctx, const void* x, void* y, const float* gamma, const float* beta, const float* pe,
                   int64_t rows, int C, int T_tok, int F, float eps, cudaStream_t st) {
  const int rows_per_block = 256 / G;
  const int64_t need = (rows + rows_per_block - 1) / rows_per_block;
  // flag 17 = 1: exactly one wave of resident blocks (3 per SM) walking the rows; = 2: the same with the next row of a lane
  // group prefetched during the arithmetic of the current one (118 registers -> 2 blocks per SM; VPL <= 5 only);
  // 0: up to 16 blocks per SM (5.3 waves at the 64 x 64 level, the last one a third full)
  if (ctx->ln_persist == 2 && VPL <= 5) {
    mmgt_launch(ctx, layernorm_grp_kernel<T, VPL, G, (VPL <= 5)>, dim3((int)std::min<int64_t>(need, (int64_t)ctx->num_sms * 2)),
                dim3(256), sizeof(float) * 2 * C, st, (const T*)x, (T*)y, gamma, beta, pe, rows, C, T_tok, F, eps);
    return;
  }
  const int64_t blocks = std::min<int64_t>(need, (int64_t)ctx->num_sms * (ctx->ln_persist ? 3 : 16));
  mmgt_launch(ctx, layernorm_grp_kernel<T, VPL, G, false>, dim3((int)blocks), dim3(256), sizeof(float) * 2 * C, st, (const T*)x,
              (T*)y, gamma, beta, pe, rows, C, T_tok, F, eps);
}

// Picks (VPL, G) with VPL * G == C / VEC for the widths of the full model (320 / 640 / 1280); false => generic kernel.
template <typename T>
bool try_ln_grp(mmgt_ctx* ctx, const void* x, void* y, const float* gamma, const float* beta, const float* pe, int64_t rows,
                int C, int T_tok, int F, float eps, cudaStream_t st) {
  const int Cv = C / VecOf<T>::N;
#define LN_CASE(VPL_, G_)                                                                         \
  if (Cv == VPL_ * G_) {                                                                          \
    launch_ln_grp<T, VPL_, G_>(ctx, x, y, gamma, beta, pe, rows, C, T_tok, F, eps, st);           \
    return true;                                                                                  \
  }
  LN_CASE(5, 8) LN_CASE(5, 16) LN_CASE(5, 32) LN_CASE(10, 32)
#undef LN_CASE
  return false;
}

// Row statistics only: (mean, rstd) per row for a LayerNorm that is fused into the GEMM consuming it (mmgt_gemm
// rowstats / colsum): one read of the tensor, 8 bytes written per row -- the normalised tensor is never materialised.
// Same two-pass (mean, centred sum of squares) arithmetic and lane layout as layernorm_grp_kernel.
template <typename T, int VPL, int G>
__global__ void __launch_bounds__(256)
row_stats_grp_kernel(const T* __restrict__ x, float2* __restrict__ stats, int64_t rows, int C, int64_t ld, float eps) {
  constexpr int VEC = VecOf<T>::N;
  pdl_prologue();
  const int gl = threadIdx.x % G;
  const int64_t grp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
  const int64_t ngrp = (int64_t)gridDim.x * blockDim.x / G;
  const float inv_c = 1.f / (float)C;
  for (int64_t row = grp; row < rows; row += ngrp) {
    float v[VPL][VEC];
    const T* xr = x + row * ld;
#pragma unroll
    for (int j = 0; j < VPL; ++j) load_vec<T>(xr + (size_t)(gl + j * G) * VEC, v[j]);
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < VPL; ++j)
#pragma unroll
      for (int e = 0; e < VEC; ++e) sum += v[j][e];
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum * inv_c;
    float sq = 0.f;
#pragma unroll
    for (int j = 0; j < VPL; ++j)
#pragma unroll
      for (int e = 0; e < VEC; ++e) { const float d = v[j][e] - mean; sq += d * d; }
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    if (gl == 0) stats[row] = make_float2(mean, rsqrtf(sq * inv_c + eps));
  }
}

// generic widths: one warp per row
template <typename T>
__global__ void __launch_bounds__(256)
row_stats_kernel(const T* __restrict__ x, float2* __restrict__ stats, int64_t rows, int C, int64_t ld, float eps) {
  constexpr int VEC = VecOf<T>::N;
  const int Cv = C / VEC, lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t row = warp; row < rows; row += nwarps) {
    float sum = 0.f;
    for (int cv = lane; cv < Cv; cv += 32) {
      float v[VEC];
      load_vec<T>(x + row * ld + (size_t)cv * VEC, v);
#pragma unroll
      for (int e = 0; e < VEC; ++e) sum += v[e];
    }
    const float mean = warp_sum(sum) / (float)C;
    float sq = 0.f;
    for (int cv = lane; cv < Cv; cv += 32) {
      float v[VEC];
      load_vec<T>(x + row * ld + (size_t)cv * VEC, v);
#pragma unroll
      for (int e = 0; e < VEC; ++e) { const float d = v[e] - mean; sq += d * d; }
    }
    sq = warp_sum(sq);
    if (lane == 0) stats[row] = make_float2(mean, rsqrtf(sq / (float)C + eps));
  }
}

template <typename T>
int launch_row_stats(mmgt_ctx* ctx, const void* x, float* stats, int64_t rows, int C, int64_t ld, float eps, cudaStream_t st) {
  const int Cv = C / VecOf<T>::N;
#define RS_CASE(VPL_, G_)                                                                                              \
  if (Cv == VPL_ * G_) {                                                                                               \
    const int rpb = 256 / G_;                                                                                          \
    const int64_t blocks = std::min<int64_t>((rows + rpb - 1) / rpb, (int64_t)ctx->num_sms * 16);                      \
    MMGT_CUDA_OK(mmgt_launch(ctx, row_stats_grp_kernel<T, VPL_, G_>, dim3((int)blocks), dim3(256), 0, st, (const T*)x, \
                             reinterpret_cast<float2*>(stats), rows, C, ld, eps));                                     \
    return 0;                                                                                                          \
  }
  RS_CASE(5, 8) RS_CASE(5, 16) RS_CASE(5, 32) RS_CASE(10, 32) RS_CASE(1, 8) RS_CASE(2, 8) RS_CASE(4, 8)
#undef RS_CASE
  const int64_t blocks = std::min<int64_t>((rows + 7) / 8, (int64_t)ctx->num_sms * 16);
  row_stats_kernel<T><<<(int)blocks, 256, 0, st>>>((const T*)x, reinterpret_cast<float2*>(stats), rows, C, ld, eps);
  return 0;
}

}  // namespace

extern "C" int mmgt_row_stats(mmgt_ctx* ctx, const void* x, float* stats, int64_t rows, int C, int64_t ld, float eps,
                              int dtype, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  MMGT_CHECK_ARG(ctx && x && stats && rows > 0 && C > 0, MMGT_E_INVALID, "row_stats: bad args");
  if (ld <= 0) ld = C;
  const int vec = dtype == MMGT_F32 ? 4 : 8;
  MMGT_CHECK_ARG(C % vec == 0 && ld % vec == 0 && ld >= C, MMGT_E_ALIGN, "row_stats: C and ld must be multiples of %d", vec);
  MMGT_CHECK_ARG(aligned16(x) && (reinterpret_cast<uintptr_t>(stats) & 7u) == 0, MMGT_E_ALIGN, "row_stats: alignment");
  int rc;
  if (dtype == MMGT_F32) rc = launch_row_stats<float>(ctx, x, stats, rows, C, ld, eps, st);
  else if (dtype == MMGT_BF16) rc = launch_row_stats<bf16>(ctx, x, stats, rows, C, ld, eps, st);
  else { mmgt_set_error("row_stats: bad dtype %d", dtype); return MMGT_E_INVALID; }
  if (rc) return rc;
  MMGT_LAUNCH_OK(ctx);
  return 0;
}

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
  const int Cv = C / vec;
  const int nv = (Cv + GN_THREADS - 1) / GN_THREADS;
  const int lanes = std::min(Cv, GN_THREADS);
  const int rpp = GN_THREADS / lanes;
  const size_t smem = sizeof(float) * (2 * C + 2 * groups);
  if (ctx->gn_split) {
    // statistics kernel + normalise kernel: grid (token chunks, frames) = ONE resident wave (3 CTAs per SM): with 8 waves'
    // worth of short CTAs the tail and the per-CTA prologue / reduction cost a quarter of the kernel
    const int want = std::max(1, (ctx->num_sms * 3) / N);
    const int unr = std::max(1, 8 / nv);
    int tok_per_block = std::max((T + want - 1) / want, rpp * unr);
    tok_per_block = (tok_per_block + rpp - 1) / rpp * rpp;
    const int chunks = (T + tok_per_block - 1) / tok_per_block;
    const size_t smem_stats = std::max(sizeof(float) * 2 * (size_t)rpp * C, smem);
    MMGT_CHECK_ARG(chunks <= 65535 * 16, MMGT_E_INVALID, "groupnorm: too many token chunks");
    MMGT_CUDA_OK(cudaMemsetAsync(stats_ws, 0, sizeof(double) * 2 * N * groups, st));
    dim3 grid(chunks, N);
#define GN2_LAUNCH(T_, NV_)                                                                                                  \
  do {                                                                                                                       \
    if (smem_stats > 48 * 1024) MMGT_CUDA_OK(mmgt_smem_optin(ctx, gn_stats_kernel<T_, NV_>, ctx->max_smem_optin));            \
    MMGT_CUDA_OK(mmgt_launch(ctx, gn_stats_kernel<T_, NV_>, grid, dim3(GN_THREADS), smem_stats, st, (const T_*)x1,           \
                             (const T_*)x2, stats_ws, T, C1, C2, groups, tok_per_block));                                    \
    (ctx)->launches++;                                                                                                       \
    MMGT_CUDA_OK(mmgt_launch(ctx, gn_apply_kernel<T_, NV_>, grid, dim3(GN_THREADS), smem, st, (const T_*)x1, (const T_*)x2,  \
                             (T_*)y, gamma, beta, (const double*)stats_ws, T, C1, C2, groups, eps, silu, tok_per_block));    \
  } while (0)
    if (dtype == MMGT_F32) { if (nv == 1) GN2_LAUNCH(float, 1); else if (nv == 2) GN2_LAUNCH(float, 2); else GN2_LAUNCH(float, 3); }
    else if (dtype == MMGT_BF16) { if (nv == 1) GN2_LAUNCH(bf16, 1); else if (nv == 2) GN2_LAUNCH(bf16, 2); else GN2_LAUNCH(bf16, 3); }
    else { mmgt_set_error("groupnorm: bad dtype %d", dtype); return MMGT_E_INVALID; }
#undef GN2_LAUNCH
    MMGT_LAUNCH_OK(ctx);
    return 0;
  }
  // co-resident blocks (the frame barrier inside the kernel relies on it)
  int per_sm = 0;
#define GN_OCC(T_, NV_) MMGT_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gn_fused_kernel<T_, NV_>, GN_THREADS, smem))
  if (dtype == MMGT_F32) { if (nv == 1) GN_OCC(float, 1); else if (nv == 2) GN_OCC(float, 2); else GN_OCC(float, 3); }
  else if (dtype == MMGT_BF16) { if (nv == 1) GN_OCC(bf16, 1); else if (nv == 2) GN_OCC(bf16, 2); else GN_OCC(bf16, 3); }
  else { mmgt_set_error("groupnorm: bad dtype %d", dtype); return MMGT_E_INVALID; }
#undef GN_OCC
  MMGT_CHECK_ARG(per_sm > 0, MMGT_E_UNSUPPORTED, "groupnorm: kernel does not fit an SM (C=%d)", C);
  const int capacity = per_sm * ctx->num_sms;
  // token chunks per frame: one wave of items when the tensor is large enough
  int chunks = (int)std::max<int64_t>(1, std::min<int64_t>((T + rpp - 1) / rpp, capacity / N));
  int tok_per_block = (T + chunks - 1) / chunks;
  tok_per_block = std::max(tok_per_block, rpp);
  chunks = (T + tok_per_block - 1) / tok_per_block;
  const int grid = std::min(N * chunks, capacity);
  // workspace: 2 * N * groups doubles of statistics, then N arrival counters
  unsigned* arrived = reinterpret_cast<unsigned*>(stats_ws + (size_t)2 * N * groups);
  MMGT_CUDA_OK(cudaMemsetAsync(stats_ws, 0, sizeof(double) * 2 * N * groups + sizeof(unsigned) * N, st));
#define GN_LAUNCH(T_, NV_)                                                                                               \
  gn_fused_kernel<T_, NV_><<<grid, GN_THREADS, smem, st>>>((const T_*)x1, (const T_*)x2, (T_*)y, gamma, beta, stats_ws,   \
                                                           arrived, N, T, C1, C2, groups, eps, silu, tok_per_block, chunks)
  if (dtype == MMGT_F32) { if (nv == 1) GN_LAUNCH(float, 1); else if (nv == 2) GN_LAUNCH(float, 2); else GN_LAUNCH(float, 3); }
  else { if (nv == 1) GN_LAUNCH(bf16, 1); else if (nv == 2) GN_LAUNCH(bf16, 2); else GN_LAUNCH(bf16, 3); }
#undef GN_LAUNCH
  MMGT_LAUNCH_OK(ctx);
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
  if (dtype == MMGT_F32 || dtype == MMGT_BF16) {
    const bool done = dtype == MMGT_F32 ? try_ln_grp<float>(ctx, x, y, gamma, beta, pe, rows, C, T, F, eps, st)
                                        : try_ln_grp<bf16>(ctx, x, y, gamma, beta, pe, rows, C, T, F, eps, st);
    if (done) {
      MMGT_LAUNCH_OK(ctx);
      return 0;
    }
  }
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
