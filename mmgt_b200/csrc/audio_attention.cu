// MM-HAA audio cross-attention, all three regions (full / face / lip) in one launch, with the motion-mask gate,
// motion_scale and the layout of the hierarchical sum fused into the epilogue (attention.py:719-767).
//
//   out[row, r*C + h*d + :] = gate_r[row] * softmax_k( q_r[row, h] . k_r[n, k, h] * scale ) v_r[n, :, h]
//   out[row, 3C + r]        = gate_r[row]                     gate_r = mask_r[row] * motion_scale[r]
//   out[row, 3C+3 .. 3C+7]  = 0
//
// so that ONE GEMM with K = 3C + 8 against [Wz_0 Wo_0 | Wz_1 Wo_1 | Wz_2 Wo_2 | Wz_r bo_r ...] finishes
// sum_r s_r * zero_conv_r(mask_r * to_out_r(attn_r)) + x: the per-region to_out / zero-conv GEMMs, the mask
// multiplies and the 3-way add of the reference never touch HBM.
//
// The operator is HBM-bound (M <= 32 audio tokens: 0.1 % of the FLOPs; reads q3 once, writes out once), so the
// MMAs (mma.sync m16n8k16, bf16 in / fp32 accumulate) only keep the arithmetic out of the way of the memory
// pipeline.  One CTA = (frame n, region r, 256 query rows); its warps take one head each; K_r / V_r of the frame
// (M x C each) stay in shared memory for the CTA's lifetime.
#include <algorithm>

#include "common.cuh"

namespace {

constexpr int AW = 8;            // warps per CTA
constexpr int ROWS_PER_CTA = 256;
constexpr int MAXM = 32;         // audio tokens per frame (keys): one 16 x 32 score tile per query tile

struct AudioArgs {
  const bf16* q3;
  const bf16* kv6;
  const float* mask[3];
  float gate_scale[3];
  bf16* out;
  int64_t ldq, ldkv, ldo;
  int N, T, M, heads, d;
  float scale_log2e;
};

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

// DK = head dim rounded up to 16.  K / V rows keep all heads side by side (row stride C + 8 elements: conflict-free
// ldmatrix, and the 8 zero pad columns absorb the DK - d overhang of the last head); the overhang of the other
// heads reads the next head's columns, which meet the zero padding of the staged Q tile (scores) or land in
// output columns that are never stored (P V).
template <int DK>
__global__ void __launch_bounds__(AW * 32)
audio_attention_mma_kernel(const AudioArgs a) {
  constexpr int DP = DK + 8;
  extern __shared__ __align__(16) uint8_t smem_u8[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = blockIdx.z, r = blockIdx.y;
  const int d = a.d, C = a.heads * d, DPc = C + 8, dv = d >> 3, Cv = C >> 3;
  bf16* Ks = reinterpret_cast<bf16*>(smem_u8);
  bf16* Vs = Ks + MAXM * DPc;
  bf16* Qs = Vs + MAXM * DPc + (size_t)warp * 16 * DP;
  // shared-memory-only prologue: zero K / V (rows >= M, pad columns) and this warp's Q staging tile
  for (int i = threadIdx.x; i < 2 * MAXM * DPc / 8; i += AW * 32) reinterpret_cast<uint4*>(Ks)[i] = make_uint4(0, 0, 0, 0);
  for (int i = lane; i < 16 * DP / 8; i += 32) reinterpret_cast<uint4*>(Qs)[i] = make_uint4(0, 0, 0, 0);
  pdl_prologue();
  __syncthreads();
  for (int i = threadIdx.x; i < a.M * Cv; i += AW * 32) {
    const int key = i / Cv, c = (i - key * Cv) * 8;
    const bf16* src = a.kv6 + ((int64_t)n * a.M + key) * a.ldkv + (int64_t)2 * r * C + c;
    *reinterpret_cast<uint4*>(Ks + key * DPc + c) = *reinterpret_cast<const uint4*>(src);
    *reinterpret_cast<uint4*>(Vs + key * DPc + c) = *reinterpret_cast<const uint4*>(src + C);
  }
  __syncthreads();

  const uint32_t q_addr = (uint32_t)__cvta_generic_to_shared(Qs), k_addr = (uint32_t)__cvta_generic_to_shared(Ks),
                 v_addr = (uint32_t)__cvta_generic_to_shared(Vs);
  const int qrow = lane >> 2, qcol = (lane & 3) * 2;   // accumulator fragment coordinates
  const float* mask = a.mask[r];
  const float gs = a.gate_scale[r];
  const int row_begin = blockIdx.x * ROWS_PER_CTA;
  const int row_end = min(a.T, row_begin + ROWS_PER_CTA);

  for (int h = warp; h < a.heads; h += AW) {
    const int hc = h * d;
    for (int t0 = row_begin; t0 < row_end; t0 += 16) {
      // ---- stage the 16 x d query tile (rows past T stay zero)
      for (int i = lane; i < 16 * dv; i += 32) {
        const int rr = i / dv, c = (i - rr * dv) * 8;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (t0 + rr < a.T) v = *reinterpret_cast<const uint4*>(a.q3 + ((int64_t)n * a.T + t0 + rr) * a.ldq + (int64_t)r * C + hc + c);
        *reinterpret_cast<uint4*>(Qs + rr * DP + c) = v;
      }
      __syncwarp();
      // ---- S = Q K^T : four 16 x 8 accumulator tiles (keys 0-7, 8-15, 16-23, 24-31)
      float s[4][4];
#pragma unroll
      for (int t = 0; t < 4; ++t) s[t][0] = s[t][1] = s[t][2] = s[t][3] = 0.f;
#pragma unroll
      for (int ks = 0; ks < DK / 16; ++ks) {
        uint32_t qa[4];
        ldsm_x4(q_addr + (((lane & 7) + ((lane >> 3) & 1) * 8) * DP + ks * 16 + (lane >> 4) * 8) * 2, qa);
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) {
          uint32_t bb[4];
          ldsm_x4(k_addr + (((lane & 7) + (lane >> 4) * 8 + 16 * kb) * DPc + hc + ks * 16 + ((lane >> 3) & 1) * 8) * 2, bb);
          mma_bf16_16816(s[2 * kb], qa, bb[0], bb[1]);
          mma_bf16_16816(s[2 * kb + 1], qa, bb[2], bb[3]);
        }
      }
      // ---- softmax over the M keys; this lane holds rows qrow (elements 0,1) and qrow + 8 (elements 2,3) of each tile
#pragma unroll
      for (int rh = 0; rh < 2; ++rh) {
        float mx = -INFINITY;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            if (8 * t + qcol + j >= a.M) s[t][2 * rh + j] = -INFINITY;
            mx = fmaxf(mx, s[t][2 * rh + j]);
          }
        }
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
        float sum = 0.f;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const float p = exp2f((s[t][2 * rh + j] - mx) * a.scale_log2e);
            s[t][2 * rh + j] = p;
            sum += p;
          }
        }
        sum += __shfl_xor_sync(0xffffffffu, sum, 1);
        sum += __shfl_xor_sync(0xffffffffu, sum, 2);
        const float inv = 1.f / sum;
#pragma unroll
        for (int t = 0; t < 4; ++t) { s[t][2 * rh] *= inv; s[t][2 * rh + 1] *= inv; }
      }
      // accumulator layout of S == A-operand layout of P: key block kb = tiles 2kb, 2kb + 1
      uint32_t pa[2][4];
#pragma unroll
      for (int kb = 0; kb < 2; ++kb) {
        pa[kb][0] = pack2(s[2 * kb][0], s[2 * kb][1]);
        pa[kb][1] = pack2(s[2 * kb][2], s[2 * kb][3]);
        pa[kb][2] = pack2(s[2 * kb + 1][0], s[2 * kb + 1][1]);
        pa[kb][3] = pack2(s[2 * kb + 1][2], s[2 * kb + 1][3]);
      }
      const int row_lo = t0 + qrow, row_hi = t0 + qrow + 8;
      const float g_lo = row_lo < a.T ? mask[(int64_t)n * a.T + row_lo] * gs : 0.f;
      const float g_hi = row_hi < a.T ? mask[(int64_t)n * a.T + row_hi] * gs : 0.f;
      __syncwarp();   // everyone is done reading Qs before it receives O
      // ---- O = gate * P V, 16 head columns per step (two n-tiles), staged in Qs
      for (int n0 = 0; n0 < d; n0 += 16) {
        float o0[4] = {0.f, 0.f, 0.f, 0.f}, o1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) {
          uint32_t vb[4];
          ldsm_x4_trans(v_addr + (((lane & 7) + ((lane >> 3) & 1) * 8 + 16 * kb) * DPc + hc + n0 + (lane >> 4) * 8) * 2, vb);
          mma_bf16_16816(o0, pa[kb], vb[0], vb[1]);
          mma_bf16_16816(o1, pa[kb], vb[2], vb[3]);
        }
        *reinterpret_cast<uint32_t*>(Qs + qrow * DP + n0 + qcol) = pack2(o0[0] * g_lo, o0[1] * g_lo);
        *reinterpret_cast<uint32_t*>(Qs + (qrow + 8) * DP + n0 + qcol) = pack2(o0[2] * g_hi, o0[3] * g_hi);
        if (n0 + 8 < d) {
          *reinterpret_cast<uint32_t*>(Qs + qrow * DP + n0 + 8 + qcol) = pack2(o1[0] * g_lo, o1[1] * g_lo);
          *reinterpret_cast<uint32_t*>(Qs + (qrow + 8) * DP + n0 + 8 + qcol) = pack2(o1[2] * g_hi, o1[3] * g_hi);
        }
      }
      __syncwarp();
      for (int i = lane; i < 16 * dv; i += 32) {
        const int rr = i / dv, c = (i - rr * dv) * 8;
        if (t0 + rr < a.T)
          *reinterpret_cast<uint4*>(a.out + ((int64_t)n * a.T + t0 + rr) * a.ldo + (int64_t)r * C + hc + c) =
              *reinterpret_cast<const uint4*>(Qs + rr * DP + c);
      }
      if (h == 0 && lane < 16 && t0 + lane < a.T) {      // the gate columns that carry mask_r * (Wz_r bo_r) through the GEMM
        bf16* tail = a.out + ((int64_t)n * a.T + t0 + lane) * a.ldo + 3 * C;
        tail[r] = __float2bfloat16_rn(mask[(int64_t)n * a.T + t0 + lane] * gs);
        if (r == 0) {
#pragma unroll
          for (int j = 3; j < 8; ++j) tail[j] = __float2bfloat16_rn(0.f);
        }
      }
      __syncwarp();
      if (d < DK) {   // the O staging never writes columns >= d, but keep the Q padding exactly zero (cheap, warp-local)
        for (int i = lane; i < 16; i += 32) *reinterpret_cast<uint4*>(Qs + i * DP + d) = make_uint4(0, 0, 0, 0);
        __syncwarp();
      }
    }
  }
}

}  // namespace

extern "C" int mmgt_audio_attention(mmgt_ctx* ctx, const mmgt_audio_attention_params* p, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  MMGT_CHECK_ARG(ctx && p, MMGT_E_INVALID, "audio_attention: null ctx/params");
  MMGT_CHECK_ARG(p->q3 && p->kv6 && p->out && p->mask[0] && p->mask[1] && p->mask[2], MMGT_E_INVALID, "audio_attention: null tensor");
  MMGT_CHECK_ARG(p->N > 0 && p->T > 0 && p->M > 0 && p->heads > 0 && p->d > 0, MMGT_E_INVALID, "audio_attention: bad sizes");
  MMGT_CHECK_ARG(p->dtype == MMGT_BF16, MMGT_E_UNSUPPORTED, "audio_attention: bf16 only (float32 runs the per-region operators)");
  MMGT_CHECK_ARG(p->M <= MAXM, MMGT_E_UNSUPPORTED, "audio_attention: M=%d audio tokens > %d", p->M, MAXM);
  MMGT_CHECK_ARG(p->d % 8 == 0 && p->d <= 160, MMGT_E_UNSUPPORTED, "audio_attention: head dim %d (need multiple of 8, <= 160)", p->d);
  const int C = p->heads * p->d;
  MMGT_CHECK_ARG(p->ldq >= 3 * C && p->ldkv >= 6 * C && p->ldo >= 3 * C + 8, MMGT_E_INVALID, "audio_attention: leading dims too small");
  MMGT_CHECK_ARG(p->ldq % 8 == 0 && p->ldkv % 8 == 0 && p->ldo % 8 == 0 && aligned16(p->q3) && aligned16(p->kv6) && aligned16(p->out),
                 MMGT_E_ALIGN, "audio_attention: rows must be 16-byte aligned");
  MMGT_CHECK_ARG(p->N <= 65535, MMGT_E_INVALID, "audio_attention: grid too large");
  AudioArgs a{};
  a.q3 = (const bf16*)p->q3; a.kv6 = (const bf16*)p->kv6; a.out = (bf16*)p->out;
  for (int r = 0; r < 3; ++r) { a.mask[r] = p->mask[r]; a.gate_scale[r] = p->scale[r]; }
  a.ldq = p->ldq; a.ldkv = p->ldkv; a.ldo = p->ldo;
  a.N = p->N; a.T = p->T; a.M = p->M; a.heads = p->heads; a.d = p->d;
  a.scale_log2e = p->softmax_scale * 1.4426950408889634f;
  const int DK = (p->d + 15) / 16 * 16;
  const size_t smem = ((size_t)2 * MAXM * (C + 8) + (size_t)AW * 16 * (DK + 8)) * 2;
  MMGT_CHECK_ARG((int)smem <= ctx->max_smem_optin, MMGT_E_UNSUPPORTED, "audio_attention: %zu B of shared memory for C=%d", smem, C);
  dim3 grid((p->T + ROWS_PER_CTA - 1) / ROWS_PER_CTA, 3, p->N);
#define ALAUNCH(DK_)                                                                                                  \
  do {                                                                                                                \
    MMGT_CUDA_OK(mmgt_smem_optin(ctx, audio_attention_mma_kernel<DK_>, ctx->max_smem_optin));          \
    MMGT_CUDA_OK(mmgt_launch(ctx, audio_attention_mma_kernel<DK_>, grid, dim3(AW * 32), smem, st, a));                \
  } while (0)
  switch (DK) {
    case 16: ALAUNCH(16); break;
    case 32: ALAUNCH(32); break;
    case 48: ALAUNCH(48); break;
    case 64: ALAUNCH(64); break;
    case 80: ALAUNCH(80); break;
    case 96: ALAUNCH(96); break;
    case 112: ALAUNCH(112); break;
    case 128: ALAUNCH(128); break;
    case 144: ALAUNCH(144); break;
    default: ALAUNCH(160); break;
  }
#undef ALAUNCH
  MMGT_LAUNCH_OK(ctx);
  return 0;
}
