// CUDA-core attention kernels (fp32 math):
//   * mmgt_attention          flash-style softmax(q k^T) v over up to two key segments  [per-frame ; shared bank]
//   * mmgt_temporal_attention motion-module attention over the F frames of every (batch, pixel) without re-layout
// They are the float32-tier path, the audio cross-attention path (Lk = 32) and the shape-generic path.
#include "common.cuh"

namespace {

constexpr int AQ = 64;        // queries per block
constexpr int AKV = 32;       // keys per tile (one per lane)
constexpr int AWARPS = 4;
constexpr int QG = 4;         // queries processed together per warp

__host__ __device__ inline int padded_stride(int d) { return ((d / 4) % 2 == 0) ? d + 4 : d + 8; }

template <typename T>
__device__ __forceinline__ void load4g(const T* p, float (&o)[4]) {
  if constexpr (sizeof(T) == 4) {
    float4 t = *reinterpret_cast<const float4*>(p);
    o[0] = t.x; o[1] = t.y; o[2] = t.z; o[3] = t.w;
  } else {
    uint2 t = *reinterpret_cast<const uint2*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
    float2 a = __bfloat1622float2(h[0]), b = __bfloat1622float2(h[1]);
    o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y;
  }
}

template <typename T, int DPL>
__global__ void __launch_bounds__(AWARPS * 32)
attention_kernel(mmgt_attention_params p) {
  pdl_prologue();
  extern __shared__ float smem[];
  const int d = p.d, S = padded_stride(d);
  float* Qs = smem;                 // [AQ][S]
  float* Ks = Qs + AQ * S;          // [AKV][S]
  float* Vs = Ks + AKV * S;         // [AKV][S]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * AQ;
  const int d4 = d >> 2;
  const T* qg = reinterpret_cast<const T*>(p.q) + ((int64_t)n * p.Lq) * p.ldq + h * d;
  const int64_t kvb = p.kv_batch_stride ? p.kv_batch_stride : (int64_t)p.Lk * p.ldk;
  const int64_t vvb = p.kv_batch_stride ? p.kv_batch_stride : (int64_t)p.Lk * p.ldv;
  const T* kg = reinterpret_cast<const T*>(p.k) + (int64_t)n * kvb + h * d;
  const T* vg = reinterpret_cast<const T*>(p.v) + (int64_t)n * vvb + h * d;
  int seg2 = -1;
  if (p.k2) seg2 = p.seg2_index ? p.seg2_index[n] : 0;
  const T* k2g = seg2 >= 0 ? reinterpret_cast<const T*>(p.k2) + ((int64_t)seg2 * p.Lk2) * p.ldk2 + h * d : nullptr;
  const T* v2g = seg2 >= 0 ? reinterpret_cast<const T*>(p.v2) + ((int64_t)seg2 * p.Lk2) * p.ldv2 + h * d : nullptr;
  const int Ltot = p.Lk + (seg2 >= 0 ? p.Lk2 : 0);

  // stage Q (pre-scaled)
  for (int i = tid; i < AQ * d4; i += AWARPS * 32) {
    int r = i / d4, c = (i % d4) * 4;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (q0 + r < p.Lq) load4g<T>(qg + (int64_t)(q0 + r) * p.ldq + c, v);
#pragma unroll
    for (int e = 0; e < 4; ++e) Qs[r * S + c + e] = v[e] * p.scale;
  }

  float o[AQ / AWARPS][DPL];
  float mrun[AQ / AWARPS], lrun[AQ / AWARPS];
#pragma unroll
  for (int i = 0; i < AQ / AWARPS; ++i) {
    mrun[i] = -INFINITY; lrun[i] = 0.f;
#pragma unroll
    for (int j = 0; j < DPL; ++j) o[i][j] = 0.f;
  }

  for (int kt = 0; kt < Ltot; kt += AKV) {
    __syncthreads();  // previous tile fully consumed (also orders the Q staging before first use)
    for (int i = tid; i < AKV * d4; i += AWARPS * 32) {
      int r = i / d4, c = (i % d4) * 4;
      int key = kt + r;
      float kv[4] = {0.f, 0.f, 0.f, 0.f}, vv[4] = {0.f, 0.f, 0.f, 0.f};
      if (key < p.Lk) {
        load4g<T>(kg + (int64_t)key * p.ldk + c, kv);
        load4g<T>(vg + (int64_t)key * p.ldv + c, vv);
      } else if (key < Ltot) {
        load4g<T>(k2g + (int64_t)(key - p.Lk) * p.ldk2 + c, kv);
        load4g<T>(v2g + (int64_t)(key - p.Lk) * p.ldv2 + c, vv);
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) { Ks[r * S + c + e] = kv[e]; Vs[r * S + c + e] = vv[e]; }
    }
    __syncthreads();
    const bool key_ok = kt + lane < Ltot;
#pragma unroll
    for (int g = 0; g < AQ / AWARPS / QG; ++g) {
      float s[QG];
#pragma unroll
      for (int qi = 0; qi < QG; ++qi) s[qi] = 0.f;
      const float* krow = Ks + lane * S;
      const float* qrow = Qs + (warp * (AQ / AWARPS) + g * QG) * S;
      for (int c = 0; c < d; c += 4) {
        float4 k4 = *reinterpret_cast<const float4*>(krow + c);
#pragma unroll
        for (int qi = 0; qi < QG; ++qi) {
          float4 q4 = *reinterpret_cast<const float4*>(qrow + qi * S + c);
          s[qi] = fmaf(q4.x, k4.x, s[qi]); s[qi] = fmaf(q4.y, k4.y, s[qi]);
          s[qi] = fmaf(q4.z, k4.z, s[qi]); s[qi] = fmaf(q4.w, k4.w, s[qi]);
        }
      }
      float pr[QG];
#pragma unroll
      for (int qi = 0; qi < QG; ++qi) {
        const int qq = g * QG + qi;
        float sv = key_ok ? s[qi] : -INFINITY;
        float mnew = fmaxf(mrun[qq], warp_max(sv));
        float corr = __expf(mrun[qq] - mnew);   // exp(-inf) = 0 on the first tile
        float pv = key_ok ? __expf(sv - mnew) : 0.f;
        lrun[qq] = lrun[qq] * corr + warp_sum(pv);
        mrun[qq] = mnew;
        pr[qi] = pv;
#pragma unroll
        for (int j = 0; j < DPL; ++j) o[qq][j] *= corr;
      }
      for (int kk = 0; kk < AKV; ++kk) {
        float vreg[DPL];
#pragma unroll
        for (int j = 0; j < DPL; ++j) {
          int c = lane + 32 * j;
          vreg[j] = c < d ? Vs[kk * S + c] : 0.f;
        }
#pragma unroll
        for (int qi = 0; qi < QG; ++qi) {
          float pk = __shfl_sync(0xffffffffu, pr[qi], kk);
#pragma unroll
          for (int j = 0; j < DPL; ++j) o[g * QG + qi][j] = fmaf(pk, vreg[j], o[g * QG + qi][j]);
        }
      }
    }
  }

  T* og = reinterpret_cast<T*>(p.out) + ((int64_t)n * p.Lq) * p.ldo + h * d;
#pragma unroll
  for (int i = 0; i < AQ / AWARPS; ++i) {
    int q = q0 + warp * (AQ / AWARPS) + i;
    if (q >= p.Lq) continue;
    float inv = 1.f / lrun[i];
#pragma unroll
    for (int j = 0; j < DPL; ++j) {
      int c = lane + 32 * j;
      if (c < d) og[(int64_t)q * p.ldo + c] = from_f32<T>(o[i][j] * inv);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Temporal attention.  Row r = (b*F + f)*T + t of qkv holds [q | k | v] (3C).  One warp owns one
// (b, t, head): lane = query frame; K, V, Q of the F frames are staged in shared memory.
constexpr int TW = 4;  // warps per block
template <typename T>
__global__ void __launch_bounds__(TW * 32)
temporal_attention_kernel(const T* __restrict__ qkv, T* __restrict__ out, int B, int F, int T_tok, int heads, int d,
                          float scale) {
  extern __shared__ float smem[];
  const int S = padded_stride(d);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* Qs = smem + (size_t)warp * 3 * F * S;
  float* Ks = Qs + F * S;
  float* Vs = Ks + F * S;
  const int C = heads * d, d4 = d >> 2;
  const int64_t total = (int64_t)B * T_tok * heads;
  for (int64_t w = (int64_t)blockIdx.x * TW + warp; w < total; w += (int64_t)gridDim.x * TW) {
    const int h = w % heads;
    const int64_t bt = w / heads;
    const int t = bt % T_tok, b = bt / T_tok;
    __syncwarp();
    for (int i = lane; i < F * d4; i += 32) {
      int f = i / d4, c = (i % d4) * 4;
      const T* row = qkv + (((int64_t)b * F + f) * T_tok + t) * (3 * C) + h * d + c;
      float qv[4], kv[4], vv[4];
      load4g<T>(row, qv); load4g<T>(row + C, kv); load4g<T>(row + 2 * C, vv);
#pragma unroll
      for (int e = 0; e < 4; ++e) { Qs[f * S + c + e] = qv[e] * scale; Ks[f * S + c + e] = kv[e]; Vs[f * S + c + e] = vv[e]; }
    }
    __syncwarp();
    float s[32];
#pragma unroll
    for (int f = 0; f < 32; ++f) s[f] = 0.f;
    const int lq = lane < F ? lane : 0;
    for (int c = 0; c < d; c += 4) {
      float4 q4 = *reinterpret_cast<const float4*>(Qs + lq * S + c);
#pragma unroll
      for (int f = 0; f < 32; ++f) {
        if (f < F) {
          float4 k4 = *reinterpret_cast<const float4*>(Ks + f * S + c);
          s[f] = fmaf(q4.x, k4.x, s[f]); s[f] = fmaf(q4.y, k4.y, s[f]);
          s[f] = fmaf(q4.z, k4.z, s[f]); s[f] = fmaf(q4.w, k4.w, s[f]);
        }
      }
    }
    float mx = -INFINITY;
#pragma unroll
    for (int f = 0; f < 32; ++f) if (f < F) mx = fmaxf(mx, s[f]);
    float sum = 0.f;
#pragma unroll
    for (int f = 0; f < 32; ++f) if (f < F) { s[f] = __expf(s[f] - mx); sum += s[f]; }
    const float inv = 1.f / sum;
    __syncwarp();  // all lanes finished reading Qs before it is reused for the output
    for (int c = 0; c < d; c += 4) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int f = 0; f < 32; ++f) {
        if (f < F) {
          float4 v4 = *reinterpret_cast<const float4*>(Vs + f * S + c);
          acc.x = fmaf(s[f], v4.x, acc.x); acc.y = fmaf(s[f], v4.y, acc.y);
          acc.z = fmaf(s[f], v4.z, acc.z); acc.w = fmaf(s[f], v4.w, acc.w);
        }
      }
      if (lane < F) *reinterpret_cast<float4*>(Qs + lane * S + c) = make_float4(acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv);
    }
    __syncwarp();
    for (int i = lane; i < F * d; i += 32) {
      int f = i / d, c = i % d;
      out[(((int64_t)b * F + f) * T_tok + t) * C + h * d + c] = from_f32<T>(Qs[f * S + c]);
    }
  }
}


// ------------------------------------------------------------------------------------------------
// bf16 temporal attention on warp-level tensor-core MMAs (m16n8k16): one warp owns one (b, t, head); the
// F <= 16 frames form one 16-row tile.  S = Q K^T (16 x 16 x d), softmax on the accumulator fragments,
// O = P V (16 x d x 16).  This operator is HBM-bound (reads 3*C, writes C per row, ~0.03 % of the FLOPs), so
// the point of the MMAs is only to get the arithmetic out of the way of the memory pipeline.
constexpr int TMW = 4;  // warps per block
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// DK = head dim rounded up to 16 (k extent of Q K^T); row stride DP = DK + 8 elements keeps ldmatrix conflict-free
template <int DK>
__global__ void __launch_bounds__(TMW * 32)
temporal_attention_mma_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ out, int B, int F, int T_tok, int heads, int d,
                              float scale_log2e) {
  pdl_prologue();
  constexpr int DP = DK + 8;
  extern __shared__ __align__(16) uint8_t smem_u8[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  bf16* Qs = reinterpret_cast<bf16*>(smem_u8) + (size_t)warp * 3 * 16 * DP;
  bf16* Ks = Qs + 16 * DP;
  bf16* Vs = Ks + 16 * DP;
  // zero once: rows >= F and columns >= d are never written again
  for (int i = lane; i < 3 * 16 * DP / 8; i += 32) reinterpret_cast<uint4*>(Qs)[i] = make_uint4(0, 0, 0, 0);
  __syncwarp();
  const int C = heads * d, dv = d >> 3;   // 16-byte vectors per row
  const int64_t total = (int64_t)B * T_tok * heads;
  const uint32_t q_addr = (uint32_t)__cvta_generic_to_shared(Qs), k_addr = (uint32_t)__cvta_generic_to_shared(Ks),
                 v_addr = (uint32_t)__cvta_generic_to_shared(Vs);
  const int qrow = lane >> 2, qcol = (lane & 3) * 2;   // accumulator fragment coordinates
  for (int64_t w = (int64_t)blockIdx.x * TMW + warp; w < total; w += (int64_t)gridDim.x * TMW) {
    const int h = w % heads;
    const int64_t bt = w / heads;
    const int t = bt % T_tok, b = bt / T_tok;
    for (int i = lane; i < F * dv; i += 32) {
      const int f = i / dv, c = (i - f * dv) * 8;
      const bf16* row = qkv + (((int64_t)b * F + f) * T_tok + t) * (3 * C) + h * d + c;
      *reinterpret_cast<uint4*>(Qs + f * DP + c) = *reinterpret_cast<const uint4*>(row);
      *reinterpret_cast<uint4*>(Ks + f * DP + c) = *reinterpret_cast<const uint4*>(row + C);
      *reinterpret_cast<uint4*>(Vs + f * DP + c) = *reinterpret_cast<const uint4*>(row + 2 * C);
    }
    __syncwarp();
    // ---- S = Q K^T : two 16x8 accumulator tiles (keys 0-7, 8-15)
    float s0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int ks = 0; ks < DK / 16; ++ks) {
      uint32_t a[4], bb[4];
      ldsm_x4(q_addr + (((lane & 7) + ((lane >> 3) & 1) * 8) * DP + ks * 16 + (lane >> 4) * 8) * 2, a);
      ldsm_x4(k_addr + (((lane & 7) + (lane >> 4) * 8) * DP + ks * 16 + ((lane >> 3) & 1) * 8) * 2, bb);
      mma_bf16_16816(s0, a, bb[0], bb[1]);
      mma_bf16_16816(s1, a, bb[2], bb[3]);
    }
    // ---- softmax over keys (columns); this lane holds rows qrow, qrow+8 and key columns qcol,+1 (+8)
    float p[8] = {s0[0], s0[1], s1[0], s1[1], s0[2], s0[3], s1[2], s1[3]};  // [row lo: k0,k1,k8,k9 | row hi: ...]
    const int kc[4] = {qcol, qcol + 1, qcol + 8, qcol + 9};
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (kc[j] >= F) p[r * 4 + j] = -INFINITY;
        mx = fmaxf(mx, p[r * 4 + j]);
      }
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
      float sum = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        p[r * 4 + j] = exp2f((p[r * 4 + j] - mx) * scale_log2e);
        sum += p[r * 4 + j];
      }
      sum += __shfl_xor_sync(0xffffffffu, sum, 1);
      sum += __shfl_xor_sync(0xffffffffu, sum, 2);
      const float inv = 1.f / sum;
#pragma unroll
      for (int j = 0; j < 4; ++j) p[r * 4 + j] *= inv;
    }
    // accumulator layout of S == A-operand layout of P (16 x 16)
    const uint32_t pa[4] = {pack2(p[0], p[1]), pack2(p[4], p[5]), pack2(p[2], p[3]), pack2(p[6], p[7])};
    __syncwarp();   // everyone is done reading Qs before it receives O
    // ---- O = P V, 16 d-columns per step (two n-tiles)
    for (int n0 = 0; n0 < d; n0 += 16) {
      uint32_t vb[4];
      ldsm_x4_trans(v_addr + (((lane & 7) + ((lane >> 3) & 1) * 8) * DP + n0 + (lane >> 4) * 8) * 2, vb);
      float o0[4] = {0.f, 0.f, 0.f, 0.f}, o1[4] = {0.f, 0.f, 0.f, 0.f};
      mma_bf16_16816(o0, pa, vb[0], vb[1]);
      mma_bf16_16816(o1, pa, vb[2], vb[3]);
      // rows qrow / qrow+8, columns n0 + qcol (+8 for the second tile) -> staged in Qs
      *reinterpret_cast<uint32_t*>(Qs + qrow * DP + n0 + qcol) = pack2(o0[0], o0[1]);
      *reinterpret_cast<uint32_t*>(Qs + (qrow + 8) * DP + n0 + qcol) = pack2(o0[2], o0[3]);
      if (n0 + 8 < d) {
        *reinterpret_cast<uint32_t*>(Qs + qrow * DP + n0 + 8 + qcol) = pack2(o1[0], o1[1]);
        *reinterpret_cast<uint32_t*>(Qs + (qrow + 8) * DP + n0 + 8 + qcol) = pack2(o1[2], o1[3]);
      }
    }
    __syncwarp();
    for (int i = lane; i < F * dv; i += 32) {
      const int f = i / dv, c = (i - f * dv) * 8;
      *reinterpret_cast<uint4*>(out + (((int64_t)b * F + f) * T_tok + t) * C + h * d + c) =
          *reinterpret_cast<const uint4*>(Qs + f * DP + c);
    }
    __syncwarp();
    // rows >= F of Qs received garbage-free zeros only if F >= 8 rows hi part unused; re-zero what O staging touched
    if (F < 16) {
      for (int i = lane; i < (16 - F) * (DP / 8); i += 32) {
        const int r = F + i / (DP / 8), c = (i % (DP / 8)) * 8;
        *reinterpret_cast<uint4*>(Qs + r * DP + c) = make_uint4(0, 0, 0, 0);
      }
    }
    if (d < DK) {   // columns [d, DK) of the rows O staging wrote (d % 16 == 8 case writes none beyond d; keep exact)
      for (int i = lane; i < F; i += 32) *reinterpret_cast<uint4*>(Qs + i * DP + d) = make_uint4(0, 0, 0, 0);
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------------
// Temporal attention, row-coalesced form (head dim <= 80): one CTA of `heads` warps owns one (batch, pixel) at a time and
// moves WHOLE token rows -- the F rows (q | k | v of all heads, 3C values, contiguous) come in through cp.async into a
// double-buffered shared-memory stage, so every global access is a full 16-byte vector of a contiguous row (the per-head
// kernel above fetches 80-byte head slices: 2.5 sectors each, one (b, t, head) per warp with nothing in flight while it
// computes).  Warp h then runs the same m16n8k16 sequence on head h and writes O over its Q slice; the F output rows
// (C values, contiguous) are stored by all threads.  The loads of pixel i + 1 are in flight during the math of pixel i.
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int DK>
__global__ void __launch_bounds__(512)
temporal_attention_rows_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ out, int B, int F, int T_tok, int heads, int d,
                               float scale_log2e) {
  pdl_prologue();
  constexpr int DP = DK + 8;
  extern __shared__ __align__(16) uint8_t smem_u8[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nthr = blockDim.x;
  const int C = heads * d, cv = C >> 3, v3 = 3 * cv;                 // 16-byte vectors per C / per q|k|v row
  const int stage_elems = 3 * heads * 16 * DP;                        // [q|k|v][head][16 rows][DP]
  bf16* stage0 = reinterpret_cast<bf16*>(smem_u8);
  // zero both stages once: rows >= F and columns >= d are never written (loads and O stores touch the valid region only)
  for (int i = threadIdx.x; i < 2 * stage_elems / 8; i += nthr) reinterpret_cast<uint4*>(stage0)[i] = make_uint4(0, 0, 0, 0);
  __syncthreads();
  const int64_t items = (int64_t)B * T_tok;
  // Source / destination offsets of this thread's 16-byte vectors do not depend on the pixel: computed once (the
  // divisions by runtime widths cost more than the copies they address), relative to the pixel's first row / the stage.
  constexpr int MAXV = 12;                                            // F * v3 / nthr <= 16 * 3 * C / (8 * 32 * heads) = 0.1875 d
  int src_off[MAXV], dst_off[MAXV];
  const int n_in = F * v3;
#pragma unroll
  for (int k = 0; k < MAXV; ++k) {
    const int i = threadIdx.x + k * nthr;
    src_off[k] = dst_off[k] = -1;
    if (i < n_in) {
      const int f = i / v3, v = i - f * v3;
      const int which = v / cv, e = (v - which * cv) * 8;             // element inside C
      const int h = e / d, c = e - h * d;
      src_off[k] = f * T_tok * 3 * C + v * 8;
      dst_off[k] = ((which * heads + h) * 16 + f) * DP + c;
    }
  }
  constexpr int MAXO = 4;
  int osrc[MAXO], odst[MAXO];
  const int n_out = F * cv;
#pragma unroll
  for (int k = 0; k < MAXO; ++k) {
    const int i = threadIdx.x + k * nthr;
    osrc[k] = odst[k] = -1;
    if (i < n_out) {
      const int f = i / cv, e = (i - f * cv) * 8;
      const int h = e / d, c = e - h * d;
      osrc[k] = (h * 16 + f) * DP + c;
      odst[k] = f * T_tok * C + e;
    }
  }
  auto issue = [&](int64_t item, int buf) {
    if (item < items) {
      const int t = item % T_tok, b = item / T_tok;
      bf16* st = stage0 + (size_t)buf * stage_elems;
      const bf16* base = qkv + ((int64_t)b * F * T_tok + t) * (3 * C);
#pragma unroll
      for (int k = 0; k < MAXV; ++k)
        if (src_off[k] >= 0) cp_async16(st + dst_off[k], base + src_off[k]);
    }
    cp_async_commit();
  };
  const int qrow = lane >> 2, qcol = (lane & 3) * 2;   // accumulator fragment coordinates
  issue(blockIdx.x, 0);
  int it = 0;
  for (int64_t item = blockIdx.x; item < items; item += gridDim.x, ++it) {
    const int buf = it & 1;
    issue(item + gridDim.x, buf ^ 1);
    cp_async_wait<1>();
    __syncthreads();
    bf16* st = stage0 + (size_t)buf * stage_elems;
    bf16* Qs = st + (size_t)warp * 16 * DP;
    bf16* Ks = st + (size_t)(heads + warp) * 16 * DP;
    bf16* Vs = st + (size_t)(2 * heads + warp) * 16 * DP;
    const uint32_t q_addr = (uint32_t)__cvta_generic_to_shared(Qs), k_addr = (uint32_t)__cvta_generic_to_shared(Ks),
                   v_addr = (uint32_t)__cvta_generic_to_shared(Vs);
    // ---- S = Q K^T : two 16x8 accumulator tiles (keys 0-7, 8-15)
    float s0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int ks = 0; ks < DK / 16; ++ks) {
      uint32_t a[4], bb[4];
      ldsm_x4(q_addr + (((lane & 7) + ((lane >> 3) & 1) * 8) * DP + ks * 16 + (lane >> 4) * 8) * 2, a);
      ldsm_x4(k_addr + (((lane & 7) + (lane >> 4) * 8) * DP + ks * 16 + ((lane >> 3) & 1) * 8) * 2, bb);
      mma_bf16_16816(s0, a, bb[0], bb[1]);
      mma_bf16_16816(s1, a, bb[2], bb[3]);
    }
    float p[8] = {s0[0], s0[1], s1[0], s1[1], s0[2], s0[3], s1[2], s1[3]};  // [row lo: k0,k1,k8,k9 | row hi: ...]
    const int kc[4] = {qcol, qcol + 1, qcol + 8, qcol + 9};
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (kc[j] >= F) p[r * 4 + j] = -INFINITY;
        mx = fmaxf(mx, p[r * 4 + j]);
      }
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
      float sum = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        p[r * 4 + j] = exp2f((p[r * 4 + j] - mx) * scale_log2e);
        sum += p[r * 4 + j];
      }
      sum += __shfl_xor_sync(0xffffffffu, sum, 1);
      sum += __shfl_xor_sync(0xffffffffu, sum, 2);
      const float inv = 1.f / sum;
#pragma unroll
      for (int j = 0; j < 4; ++j) p[r * 4 + j] *= inv;
    }
    const uint32_t pa[4] = {pack2(p[0], p[1]), pack2(p[4], p[5]), pack2(p[2], p[3]), pack2(p[6], p[7])};
    __syncwarp();   // every lane is done reading this head's Q before it receives O
    for (int n0 = 0; n0 < d; n0 += 16) {
      uint32_t vb[4];
      ldsm_x4_trans(v_addr + (((lane & 7) + ((lane >> 3) & 1) * 8) * DP + n0 + (lane >> 4) * 8) * 2, vb);
      float o0[4] = {0.f, 0.f, 0.f, 0.f}, o1[4] = {0.f, 0.f, 0.f, 0.f};
      mma_bf16_16816(o0, pa, vb[0], vb[1]);
      mma_bf16_16816(o1, pa, vb[2], vb[3]);
      // valid rows / columns only: the zero padding of the stage must survive
      if (qrow < F) *reinterpret_cast<uint32_t*>(Qs + qrow * DP + n0 + qcol) = pack2(o0[0], o0[1]);
      if (qrow + 8 < F) *reinterpret_cast<uint32_t*>(Qs + (qrow + 8) * DP + n0 + qcol) = pack2(o0[2], o0[3]);
      if (n0 + 8 < d) {
        if (qrow < F) *reinterpret_cast<uint32_t*>(Qs + qrow * DP + n0 + 8 + qcol) = pack2(o1[0], o1[1]);
        if (qrow + 8 < F) *reinterpret_cast<uint32_t*>(Qs + (qrow + 8) * DP + n0 + 8 + qcol) = pack2(o1[2], o1[3]);
      }
    }
    __syncthreads();
    {   // the F output rows of this pixel, C contiguous values each
      const int t = item % T_tok, b = item / T_tok;
      bf16* obase = out + ((int64_t)b * F * T_tok + t) * C;
#pragma unroll
      for (int k = 0; k < MAXO; ++k)
        if (osrc[k] >= 0) *reinterpret_cast<uint4*>(obase + odst[k]) = *reinterpret_cast<const uint4*>(st + osrc[k]);
    }
    __syncthreads();   // the stage is free for the loads of item + 2 * gridDim.x (issued at the top of the next iteration)
  }
  cp_async_wait<0>();
}

}  // namespace

extern "C" int mmgt_attention(mmgt_ctx* ctx, const mmgt_attention_params* p, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  MMGT_CHECK_ARG(ctx && p, MMGT_E_INVALID, "attention: null ctx/params");
  MMGT_CHECK_ARG(p->q && p->k && p->v && p->out && p->N > 0 && p->Lq > 0 && p->Lk > 0 && p->heads > 0 && p->d > 0,
                 MMGT_E_INVALID, "attention: bad args");
  MMGT_CHECK_ARG(!p->k2 || (p->v2 && p->Lk2 > 0), MMGT_E_INVALID, "attention: second segment incomplete");
  if (ctx->use_tc && mmgt_attention_tc_supported(ctx, p)) return mmgt_attention_tc(ctx, p, st);
  if (p->dtype == MMGT_BF16) MMGT_SIMT_FALLBACK(ctx, "attention");
  MMGT_CHECK_ARG(p->d % 4 == 0 && p->d <= 160, MMGT_E_UNSUPPORTED, "attention: head dim %d (need multiple of 4, <= 160)", p->d);
  const int ev = p->dtype == MMGT_F32 ? 16 : 8;
  MMGT_CHECK_ARG(p->ldq % 4 == 0 && p->ldk % 4 == 0 && p->ldv % 4 == 0 && (!p->k2 || (p->ldk2 % 4 == 0 && p->ldv2 % 4 == 0)) &&
                     p->kv_batch_stride % 4 == 0,
                 MMGT_E_ALIGN, "attention: leading dims must be multiples of 4 elements");
  MMGT_CHECK_ARG((uintptr_t)p->q % ev == 0 && (uintptr_t)p->k % ev == 0 && (uintptr_t)p->v % ev == 0 &&
                     (!p->k2 || ((uintptr_t)p->k2 % ev == 0 && (uintptr_t)p->v2 % ev == 0)),
                 MMGT_E_ALIGN, "attention: pointers must be aligned to 4 elements");
  MMGT_CHECK_ARG(p->N <= 65535 && p->heads <= 65535, MMGT_E_INVALID, "attention: grid too large");
  const int S = padded_stride(p->d);
  const size_t smem = sizeof(float) * (size_t)(AQ + 2 * AKV) * S;
  dim3 grid((p->Lq + AQ - 1) / AQ, p->heads, p->N);
  const int dpl = (p->d + 31) / 32;
#define LAUNCH(TT, DPL)                                                                                            \
  do {                                                                                                             \
    MMGT_CUDA_OK(mmgt_smem_optin(ctx, attention_kernel<TT, DPL>, ctx->max_smem_optin));          \
    MMGT_CUDA_OK(mmgt_launch(ctx, attention_kernel<TT, DPL>, grid, dim3(AWARPS * 32), smem, st, *p));            \
  } while (0)
  MMGT_DISPATCH_DTYPE(p->dtype, T_, {
    if (dpl <= 1) LAUNCH(T_, 1);
    else if (dpl == 2) LAUNCH(T_, 2);
    else if (dpl == 3) LAUNCH(T_, 3);
    else LAUNCH(T_, 5);
  });
#undef LAUNCH
  MMGT_LAUNCH_OK(ctx);
  return 0;
}

extern "C" int mmgt_temporal_attention(mmgt_ctx* ctx, const void* qkv, void* out, int B, int F, int T, int heads, int d,
                                       float scale, int dtype, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  MMGT_CHECK_ARG(ctx && qkv && out && B > 0 && F > 0 && T > 0 && heads > 0 && d > 0, MMGT_E_INVALID,
                 "temporal_attention: bad args");
  MMGT_CHECK_ARG(F <= 32, MMGT_E_UNSUPPORTED, "temporal_attention: F=%d > 32 (positional table max_len is 32)", F);
  MMGT_CHECK_ARG(d % 4 == 0, MMGT_E_UNSUPPORTED, "temporal_attention: head dim must be a multiple of 4");
  MMGT_CHECK_ARG(aligned16(qkv), MMGT_E_ALIGN, "temporal_attention: qkv must be 16B aligned");
  if (dtype == MMGT_BF16 && ctx->use_tc && ctx->temporal_rows && F <= 16 && d % 8 == 0 && d <= 80 && heads <= 16) {
    const int DK = (d + 15) / 16 * 16;
    const size_t smem = (size_t)2 * 3 * heads * 16 * (DK + 8) * 2;
    const int nthr_rows = heads * 32, cvec = heads * d / 8;
    const bool tables_ok = (F * 3 * cvec + nthr_rows - 1) / nthr_rows <= 12 && (F * cvec + nthr_rows - 1) / nthr_rows <= 4 &&
                           (int64_t)F * T * 3 * heads * d < (1ll << 31);
    if ((int)smem <= ctx->max_smem_optin && tables_ok) {
      const int per_sm = std::max(1, (int)((size_t)ctx->max_smem_optin / (smem + 1024)));
      const int64_t items = (int64_t)B * T;
      const int blocks = (int)std::min<int64_t>(items, (int64_t)ctx->num_sms * per_sm);
      const float sl2 = scale * 1.4426950408889634f;
#define RLAUNCH(DK_)                                                                                                       \
  do {                                                                                                                     \
    MMGT_CUDA_OK(mmgt_smem_optin(ctx, temporal_attention_rows_kernel<DK_>, ctx->max_smem_optin));                          \
    MMGT_CUDA_OK(mmgt_launch(ctx, temporal_attention_rows_kernel<DK_>, dim3(blocks), dim3(heads * 32), smem, st,           \
                             (const bf16*)qkv, (bf16*)out, B, F, T, heads, d, sl2));                                      \
  } while (0)
      switch (DK) {
        case 16: RLAUNCH(16); break;
        case 32: RLAUNCH(32); break;
        case 48: RLAUNCH(48); break;
        case 64: RLAUNCH(64); break;
        default: RLAUNCH(80); break;
      }
#undef RLAUNCH
      MMGT_LAUNCH_OK(ctx);
      return 0;
    }
  }
  if (dtype == MMGT_BF16 && ctx->use_tc && F <= 16 && d % 8 == 0 && d <= 160) {
    const int DK = (d + 15) / 16 * 16;
    const size_t smem = (size_t)TMW * 3 * 16 * (DK + 8) * 2;
    const int64_t total = (int64_t)B * T * heads;
    int blocks = (int)std::min<int64_t>((total + TMW - 1) / TMW, (int64_t)ctx->num_sms * 32);
    const float sl2 = scale * 1.4426950408889634f;
#define TLAUNCH(DK_)                                                                                                   \
  do {                                                                                                                 \
    MMGT_CUDA_OK(mmgt_smem_optin(ctx, temporal_attention_mma_kernel<DK_>, ctx->max_smem_optin));          \
    MMGT_CUDA_OK(mmgt_launch(ctx, temporal_attention_mma_kernel<DK_>, dim3(blocks), dim3(TMW * 32), smem, st,         \
                             (const bf16*)qkv, (bf16*)out, B, F, T, heads, d, sl2));                                   \
  } while (0)
    switch (DK) {
      case 16: TLAUNCH(16); break;
      case 32: TLAUNCH(32); break;
      case 48: TLAUNCH(48); break;
      case 64: TLAUNCH(64); break;
      case 80: TLAUNCH(80); break;
      case 96: TLAUNCH(96); break;
      case 112: TLAUNCH(112); break;
      case 128: TLAUNCH(128); break;
      case 144: TLAUNCH(144); break;
      default: TLAUNCH(160); break;
    }
#undef TLAUNCH
    MMGT_LAUNCH_OK(ctx);
    return 0;
  }
  if (dtype == MMGT_BF16) MMGT_SIMT_FALLBACK(ctx, "temporal_attention");
  const int S = padded_stride(d);
  const size_t smem = sizeof(float) * (size_t)TW * 3 * F * S;
  MMGT_CHECK_ARG((int)smem <= ctx->max_smem_optin, MMGT_E_UNSUPPORTED, "temporal_attention: smem %zu too large", smem);
  const int64_t total = (int64_t)B * T * heads;
  int blocks = (int)std::min<int64_t>((total + TW - 1) / TW, (int64_t)ctx->num_sms * 32);
  MMGT_DISPATCH_DTYPE(dtype, T_, {
    MMGT_CUDA_OK(mmgt_smem_optin(ctx, temporal_attention_kernel<T_>, ctx->max_smem_optin));
    temporal_attention_kernel<T_><<<blocks, TW * 32, smem, st>>>((const T_*)qkv, (T_*)out, B, F, T, heads, d, scale);
  });
  MMGT_LAUNCH_OK(ctx);
  return 0;
}
