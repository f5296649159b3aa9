// CUDA-core GEMM / implicit 3x3 convolution with the fused epilogue.
// This is the float32 ("1e-4 parity tier") path and the shape-generic path for operands the tcgen05 kernels do
// not take (Cin = 4 conv_in, Cout = 4 conv_out, M = batch-sized vectors).  fp32 accumulation throughout.
#include "common.cuh"

namespace {

constexpr int BM = 64, BN = 64, BK = 16, NT = 256;

struct ConvGeom {
  int H, W, Cin, stride, up, Ho, Wo;  // H, W are the stored input dims (before the optional x2 upsample)
};

struct Epi {
  const float* bias;
  const float* rowscale;
  const float* rowbias;
  const void* residual;
  int64_t ldr;
  int rows_per_group;
  float alpha;
  int64_t ld_rowbias;      // 0 = N_out
  int rowbias_mod;
  const float* rowstats;   // fused LayerNorm: (M,2) [mean, rstd]
  const float* colsum;
  int act;
};

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == 1) return silu_f(v);
  if (act == 2) return fmaxf(v, 0.f);
  return v;
}

template <typename T>
__device__ __forceinline__ void load4(const T* p, bool vec_ok, int valid, float (&o)[4]) {
  // loads up to `valid` (<=4) consecutive elements, zero-filling the rest
  if (vec_ok && valid == 4) {
    if constexpr (sizeof(T) == 4) {
      float4 t = *reinterpret_cast<const float4*>(p);
      o[0] = t.x; o[1] = t.y; o[2] = t.z; o[3] = t.w;
    } else {
      uint2 t = *reinterpret_cast<const uint2*>(p);
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
      float2 a = __bfloat1622float2(h[0]), b = __bfloat1622float2(h[1]);
      o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) o[i] = i < valid ? to_f32(p[i]) : 0.f;
  }
}

// GEGLU: logical output column j pairs W rows value = (j/gb)*2gb + j%gb and gate = value + gb
template <typename T, typename TD, bool CONV, bool GEGLU>
__global__ void __launch_bounds__(NT)
gemm_simt_kernel(const T* __restrict__ A, const T* __restrict__ Wt, TD* __restrict__ D, int M, int N_out, int K,
                 int64_t lda, int64_t ldw, int64_t ldd, Epi ep, int gb, ConvGeom cg, int a_vec_ok, int w_vec_ok) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[GEGLU ? 2 : 1][BK][BN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int lrow = tid >> 2, lk = (tid & 3) * 4;  // each thread stages 4 consecutive k of one row
  const int tx = tid & 15, ty = tid >> 4;

  // A row geometry (fixed per thread)
  const int am = m0 + lrow;
  const bool am_ok = am < M;
  int an = 0, aoh = 0, aow = 0;
  if (CONV && am_ok) {
    an = am / (cg.Ho * cg.Wo);
    int rem = am - an * cg.Ho * cg.Wo;
    aoh = rem / cg.Wo;
    aow = rem - aoh * cg.Wo;
  }
  const int Hi = CONV ? (cg.up ? 2 * cg.H : cg.H) : 0, Wi = CONV ? (cg.up ? 2 * cg.W : cg.W) : 0;
  // W rows (fixed per thread)
  const int wj = n0 + lrow;
  const bool wj_ok = wj < N_out;
  const int64_t wrow_val = GEGLU ? ((int64_t)(wj / gb) * 2 * gb + wj % gb) : wj;

  float acc[4][4], accg[GEGLU ? 4 : 1][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) { acc[i][j] = 0.f; if (GEGLU) accg[i][j] = 0.f; }

  for (int k0 = 0; k0 < K; k0 += BK) {
    const int k = k0 + lk;
    const int valid = max(0, min(4, K - k));
    float a4[4] = {0.f, 0.f, 0.f, 0.f};
    if (am_ok && valid > 0) {
      if (CONV) {
        const int tap = k / cg.Cin, c = k - tap * cg.Cin;  // Cin % 4 == 0 is enforced by the host for CONV
        int ih = aoh * cg.stride + tap / 3 - 1, iw = aow * cg.stride + tap % 3 - 1;
        if (ih >= 0 && ih < Hi && iw >= 0 && iw < Wi) {
          if (cg.up) { ih >>= 1; iw >>= 1; }
          load4<T>(A + (((int64_t)an * cg.H + ih) * cg.W + iw) * cg.Cin + c, a_vec_ok, valid, a4);
        }
      } else {
        load4<T>(A + (int64_t)am * lda + k, a_vec_ok, valid, a4);
      }
    }
    float b4[4] = {0.f, 0.f, 0.f, 0.f}, g4[4] = {0.f, 0.f, 0.f, 0.f};
    if (wj_ok && valid > 0) {
      load4<T>(Wt + wrow_val * ldw + k, w_vec_ok, valid, b4);
      if (GEGLU) load4<T>(Wt + (wrow_val + gb) * ldw + k, w_vec_ok, valid, g4);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      As[lk + i][lrow] = a4[i];
      Bs[0][lk + i][lrow] = b4[i];
      if (GEGLU) Bs[GEGLU ? 1 : 0][lk + i][lrow] = g4[i];
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float4 av = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      float4 bv = *reinterpret_cast<const float4*>(&Bs[0][kk][tx * 4]);
      const float a[4] = {av.x, av.y, av.z, av.w}, b[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      if (GEGLU) {
        float4 gv = *reinterpret_cast<const float4*>(&Bs[GEGLU ? 1 : 0][kk][tx * 4]);
        const float g[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) accg[i][j] = fmaf(a[i], g[j], accg[i][j]);
      }
    }
    __syncthreads();
  }

  const T* res = reinterpret_cast<const T*>(ep.residual);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
    const float rs = ep.rowscale ? ep.rowscale[m] : 1.f;
    int grp = ep.rowbias ? m / ep.rows_per_group : 0;
    if (ep.rowbias_mod > 0) grp %= ep.rowbias_mod;
    const int64_t ldrb = ep.ld_rowbias ? ep.ld_rowbias : N_out;
    const float mean = ep.rowstats ? ep.rowstats[2 * (int64_t)m] : 0.f, rstd = ep.rowstats ? ep.rowstats[2 * (int64_t)m + 1] : 1.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N_out) continue;
      float v;
      if (GEGLU) {
        const int64_t wv = (int64_t)(n / gb) * 2 * gb + n % gb;
        float val = acc[i][j], gate = accg[i][j];
        if (ep.rowstats) {
          val = rstd * (val - mean * ep.colsum[wv]);
          gate = rstd * (gate - mean * ep.colsum[wv + gb]);
        }
        val += ep.bias ? ep.bias[wv] : 0.f;
        gate += ep.bias ? ep.bias[wv + gb] : 0.f;
        v = val * gelu_erf_f(gate);
      } else {
        v = acc[i][j];
        if (ep.rowstats) v = rstd * (v - mean * ep.colsum[n]);
        v += ep.bias ? ep.bias[n] : 0.f;
      }
      v *= rs * ep.alpha;
      if (ep.rowbias) v += ep.rowbias[grp * ldrb + n];
      v = apply_act(v, ep.act);
      if (res) v += to_f32(res[(int64_t)m * ep.ldr + n]);
      D[(int64_t)m * ldd + n] = from_f32<TD>(v);
    }
  }
}

// Batch-sized float32 vectors (time embedding, per-resnet time projections, CLIP value/out vectors): M <= 8 rows.
// One warp per output column, lanes stride over K with float4 loads, warp-shuffle reduction.
constexpr int GEMV_MAX_M = 8;
__global__ void __launch_bounds__(256)
gemv_f32_kernel(const float* __restrict__ A, const float* __restrict__ W, float* __restrict__ D, const float* __restrict__ bias,
                int M, int N, int K, int64_t lda, int64_t ldw, int64_t ldd, float alpha) {
  const int lane = threadIdx.x & 31;
  const int warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int nwarps = gridDim.x * (blockDim.x >> 5);
  for (int n = warp; n < N; n += nwarps) {
    float acc[GEMV_MAX_M];
#pragma unroll
    for (int m = 0; m < GEMV_MAX_M; ++m) acc[m] = 0.f;
    const float* w = W + (int64_t)n * ldw;
    for (int k = lane * 4; k < K; k += 128) {
      const float4 wv = *reinterpret_cast<const float4*>(w + k);
#pragma unroll
      for (int m = 0; m < GEMV_MAX_M; ++m) {
        if (m < M) {
          const float4 av = *reinterpret_cast<const float4*>(A + (int64_t)m * lda + k);
          acc[m] = fmaf(av.x, wv.x, acc[m]); acc[m] = fmaf(av.y, wv.y, acc[m]);
          acc[m] = fmaf(av.z, wv.z, acc[m]); acc[m] = fmaf(av.w, wv.w, acc[m]);
        }
      }
    }
#pragma unroll
    for (int m = 0; m < GEMV_MAX_M; ++m) {
      if (m < M) {
        float v = warp_sum(acc[m]);
        if (lane == 0) D[(int64_t)m * ldd + n] = alpha * (v + (bias ? bias[n] : 0.f));
      }
    }
  }
}

template <typename T, typename TD, bool CONV>
int launch(mmgt_ctx* ctx, const void* A, const void* W, void* D, int M, int N, int K, int64_t lda, int64_t ldw,
           int64_t ldd, const Epi& ep, int gb, const ConvGeom& cg, cudaStream_t st) {
  const int N_out = gb ? N / 2 : N;
  dim3 grid((N_out + BN - 1) / BN, (M + BM - 1) / BM);
  MMGT_CHECK_ARG(grid.y <= 65535, MMGT_E_INVALID, "gemm: M=%d too large for the CUDA-core kernel", M);
  const int vb = sizeof(T) == 4 ? 16 : 8;  // bytes of a 4-element vector
  int a_vec = (reinterpret_cast<uintptr_t>(A) % vb == 0) && (CONV ? (cg.Cin % 4 == 0) : (lda % 4 == 0));
  int w_vec = (reinterpret_cast<uintptr_t>(W) % vb == 0) && (ldw % 4 == 0);
  if (gb)
    gemm_simt_kernel<T, TD, CONV, true><<<grid, NT, 0, st>>>((const T*)A, (const T*)W, (TD*)D, M, N_out, K, lda, ldw, ldd, ep, gb, cg, a_vec, w_vec);
  else
    gemm_simt_kernel<T, TD, CONV, false><<<grid, NT, 0, st>>>((const T*)A, (const T*)W, (TD*)D, M, N_out, K, lda, ldw, ldd, ep, gb, cg, a_vec, w_vec);
  MMGT_LAUNCH_OK(ctx);
  return 0;
}

}  // namespace

extern "C" int mmgt_gemm(mmgt_ctx* ctx, const mmgt_gemm_params* p, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  MMGT_CHECK_ARG(ctx && p, MMGT_E_INVALID, "gemm: null ctx/params");
  MMGT_CHECK_ARG(p->A && p->W && (p->D || p->exchange) && p->M > 0 && p->N > 0 && p->K > 0, MMGT_E_INVALID, "gemm: bad args");
  MMGT_CHECK_ARG(p->lda >= p->K && p->ldw >= p->K, MMGT_E_INVALID, "gemm: leading dims smaller than K");
  MMGT_CHECK_ARG(!p->rowbias || p->rows_per_group > 0, MMGT_E_INVALID, "gemm: rowbias needs rows_per_group");
  MMGT_CHECK_ARG(!p->residual || p->ldr > 0, MMGT_E_INVALID, "gemm: residual needs ldr");
  if (p->geglu_block) {
    MMGT_CHECK_ARG(p->geglu_block > 0 && p->N % (2 * p->geglu_block) == 0, MMGT_E_INVALID,
                   "gemm: N=%d not a multiple of 2*geglu_block=%d", p->N, 2 * p->geglu_block);
  }
  MMGT_CHECK_ARG(p->exchange || p->ldd >= (p->geglu_block ? p->N / 2 : p->N), MMGT_E_INVALID, "gemm: ldd too small");
  MMGT_CHECK_ARG(!p->rowstats || p->colsum, MMGT_E_INVALID, "gemm: rowstats (fused LayerNorm) needs colsum");
  MMGT_CHECK_ARG(p->act >= 0 && p->act <= 2, MMGT_E_INVALID, "gemm: act must be 0 (none), 1 (SiLU) or 2 (ReLU)");
  if (p->dtype == MMGT_BF16 && !p->out_f32 && ctx->use_tc && mmgt_gemm_tc_supported(ctx, p)) return mmgt_gemm_tc(ctx, p, st);
  // batch-sized vectors (time embedding, CLIP: M <= GEMV_MAX_M rows) run on CUDA cores by design; anything larger in
  // bf16 is a shape the tensor-core kernels do not cover
  if (p->dtype == MMGT_BF16 && !p->out_f32 && p->M > GEMV_MAX_M) MMGT_SIMT_FALLBACK(ctx, "gemm");
  MMGT_CHECK_ARG(!p->exchange, MMGT_E_UNSUPPORTED,
                 "gemm: the fused row exchange needs the bf16 tensor-core path (use mmgt_row_exchange_copy otherwise)");
  if (p->dtype == MMGT_F32 && p->M <= GEMV_MAX_M && !p->geglu_block && !p->rowscale && !p->rowbias && !p->residual &&
      !p->rowstats && !p->act &&
      p->K % 4 == 0 && p->lda % 4 == 0 && p->ldw % 4 == 0 && aligned16(p->A) && aligned16(p->W)) {
    int blocks = std::min((p->N + 7) / 8, ctx->num_sms * 8);
    gemv_f32_kernel<<<blocks, 256, 0, st>>>((const float*)p->A, (const float*)p->W, (float*)p->D, p->bias, p->M, p->N, p->K,
                                           p->lda, p->ldw, p->ldd, p->alpha);
    MMGT_LAUNCH_OK(ctx);
    return 0;
  }
  Epi ep{p->bias, p->rowscale, p->rowbias, p->residual, p->ldr, p->rows_per_group, p->alpha,
         p->ld_rowbias, p->rowbias_mod, p->rowstats, p->colsum, p->act};
  ConvGeom cg{};
  if (p->dtype == MMGT_F32) return launch<float, float, false>(ctx, p->A, p->W, p->D, p->M, p->N, p->K, p->lda, p->ldw, p->ldd, ep, p->geglu_block, cg, st);
  if (p->dtype == MMGT_BF16) {
    if (p->out_f32) return launch<bf16, float, false>(ctx, p->A, p->W, p->D, p->M, p->N, p->K, p->lda, p->ldw, p->ldd, ep, p->geglu_block, cg, st);
    return launch<bf16, bf16, false>(ctx, p->A, p->W, p->D, p->M, p->N, p->K, p->lda, p->ldw, p->ldd, ep, p->geglu_block, cg, st);
  }
  mmgt_set_error("gemm: bad dtype %d", p->dtype);
  return MMGT_E_INVALID;
}

static int conv_out_dims(const mmgt_conv3x3_params* p, int* Ho, int* Wo) {
  const int Hi = p->upsample2x ? 2 * p->H : p->H, Wi = p->upsample2x ? 2 * p->W : p->W;
  *Ho = (Hi - 1) / p->stride + 1;
  *Wo = (Wi - 1) / p->stride + 1;
  return 0;
}

extern "C" int64_t mmgt_conv3x3_workspace_bytes(mmgt_ctx* ctx, const mmgt_conv3x3_params* p) {
  if (!ctx || !p) return MMGT_E_INVALID;
  if (p->dtype == MMGT_BF16 && ctx->use_tc && ctx->conv_implicit_all && mmgt_conv3x3_tc_supported(ctx, p)) return 0;
  if (p->dtype == MMGT_BF16 && ctx->use_tc && (p->stride != 1 || p->upsample2x) && p->Cin % 64 == 0 && p->Cout % 8 == 0) {
    int Ho, Wo;
    conv_out_dims(p, &Ho, &Wo);
    return (int64_t)p->N * Ho * Wo * 9 * p->Cin * 2;
  }
  return 0;
}

extern "C" int mmgt_conv3x3(mmgt_ctx* ctx, const mmgt_conv3x3_params* p, void* workspace, int64_t workspace_bytes,
                            void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  MMGT_CHECK_ARG(ctx && p, MMGT_E_INVALID, "conv3x3: null ctx/params");
  MMGT_CHECK_ARG(p->x && p->w && p->y && p->N > 0 && p->H > 0 && p->W > 0 && p->Cin > 0 && p->Cout > 0, MMGT_E_INVALID,
                 "conv3x3: bad args");
  MMGT_CHECK_ARG(p->stride == 1 || p->stride == 2, MMGT_E_INVALID, "conv3x3: stride must be 1 or 2");
  MMGT_CHECK_ARG(!p->rowbias || p->frames_per_group > 0, MMGT_E_INVALID, "conv3x3: rowbias needs frames_per_group");
  MMGT_CHECK_ARG(p->act >= 0 && p->act <= 2, MMGT_E_INVALID, "conv3x3: act must be 0 (none), 1 (SiLU) or 2 (ReLU)");
  int Ho, Wo;
  conv_out_dims(p, &Ho, &Wo);
  const int M = p->N * Ho * Wo, K = 9 * p->Cin;
  if (p->dtype == MMGT_BF16 && ctx->use_tc) {
    if (ctx->conv_implicit_all ? mmgt_conv3x3_tc_supported(ctx, p)
                               : (p->stride == 1 && !p->upsample2x && mmgt_conv3x3_tc_supported(ctx, p)))
      return mmgt_conv3x3_tc(ctx, p, st);
    const int64_t need = mmgt_conv3x3_workspace_bytes(ctx, p);
    if (need > 0) {
      // stride-2 / upsampling convs: stage an im2col matrix, then the tensor-core GEMM with the same epilogue
      mmgt_gemm_params g{};
      g.A = workspace; g.W = p->w; g.D = p->y; g.bias = p->bias; g.rowbias = p->rowbias; g.residual = p->residual;
      g.lda = K; g.ldw = K; g.ldd = p->Cout; g.ldr = p->Cout; g.M = M; g.N = p->Cout; g.K = K;
      g.rows_per_group = p->frames_per_group * Ho * Wo; g.alpha = 1.f; g.dtype = MMGT_BF16;
      g.ld_rowbias = p->ld_rowbias; g.act = p->act;
      if (mmgt_gemm_tc_supported(ctx, &g)) {
        MMGT_CHECK_ARG(workspace && workspace_bytes >= need, MMGT_E_INVALID, "conv3x3: workspace too small (%lld < %lld)",
                       (long long)workspace_bytes, (long long)need);
        int rc = mmgt_im2col3x3(ctx, p->x, workspace, p->N, p->H, p->W, p->Cin, p->stride, p->upsample2x, MMGT_BF16, stream);
        if (rc) return rc;
        return mmgt_gemm_tc(ctx, &g, st);
      }
    }
  }
  MMGT_CHECK_ARG(p->Cin % 4 == 0, MMGT_E_UNSUPPORTED, "conv3x3: Cin must be a multiple of 4");
  if (p->dtype == MMGT_BF16) MMGT_SIMT_FALLBACK(ctx, "conv3x3");
  Epi ep{p->bias, nullptr, p->rowbias, p->residual, p->Cout, p->frames_per_group * Ho * Wo, 1.f,
         p->ld_rowbias, 0, nullptr, nullptr, p->act};
  ConvGeom cg{p->H, p->W, p->Cin, p->stride, p->upsample2x, Ho, Wo};
  if (p->dtype == MMGT_F32) return launch<float, float, true>(ctx, p->x, p->w, p->y, M, p->Cout, K, 0, K, p->Cout, ep, 0, cg, st);
  if (p->dtype == MMGT_BF16) return launch<bf16, bf16, true>(ctx, p->x, p->w, p->y, M, p->Cout, K, 0, K, p->Cout, ep, 0, cg, st);
  mmgt_set_error("conv3x3: bad dtype %d", p->dtype);
  return MMGT_E_INVALID;
}
