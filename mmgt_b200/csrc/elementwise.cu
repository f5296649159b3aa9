// Memory-bound helper kernels: layout changes, upsample / im2col, time embedding, CFG + DDIM update.
#include "common.cuh"

// ---------------------------------------------------------------- NCFHW <-> tokens (tiled transpose)
// For each (b, f): a (C x HW) matrix in src becomes (HW x C) in dst.
// tok_ld = channel stride of the token tensor (>= C; > C when the producer padded its output channels)
template <typename TS, typename TD, bool kToTokens>
__global__ void ncfhw_tokens_kernel(const TS* __restrict__ src, const TS* __restrict__ add, TD* __restrict__ dst,
                                    int B, int C, int F, int HW, int tok_ld) {
  __shared__ float tile[32][33];
  const int bf = blockIdx.z;
  const int b = bf / F, f = bf % F;
  const int c0 = blockIdx.y * 32, t0 = blockIdx.x * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
  if (kToTokens) {
#pragma unroll
    for (int i = 0; i < 32; i += 8) {
      int c = c0 + ty + i, t = t0 + tx;
      if (c < C && t < HW) {
        size_t idx = (((size_t)b * C + c) * F + f) * HW + t;
        float v = to_f32(src[idx]);
        if (add) v += to_f32(add[idx]);
        tile[ty + i][tx] = v;
      }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 32; i += 8) {
      int t = t0 + ty + i, c = c0 + tx;
      if (c < C && t < HW) dst[((size_t)bf * HW + t) * tok_ld + c] = from_f32<TD>(tile[tx][ty + i]);
    }
  } else {
#pragma unroll
    for (int i = 0; i < 32; i += 8) {
      int t = t0 + ty + i, c = c0 + tx;
      if (c < C && t < HW) tile[ty + i][tx] = to_f32(src[((size_t)bf * HW + t) * tok_ld + c]);
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 32; i += 8) {
      int c = c0 + ty + i, t = t0 + tx;
      if (c < C && t < HW) dst[(((size_t)b * C + c) * F + f) * HW + t] = from_f32<TD>(tile[tx][ty + i]);
    }
  }
}

template <bool kToTokens>
static int launch_layout(mmgt_ctx* ctx, const void* src, const void* add, void* dst, int B, int C, int F, int H,
                         int W, int sdt, int ddt, int tok_ld, cudaStream_t st) {
  MMGT_CHECK_ARG(src && dst && B > 0 && C > 0 && F > 0 && H > 0 && W > 0, MMGT_E_INVALID, "layout: bad args");
  MMGT_CHECK_ARG((size_t)B * F <= 65535, MMGT_E_INVALID, "layout: B*F too large");
  if (tok_ld <= 0) tok_ld = C;
  MMGT_CHECK_ARG(tok_ld >= C, MMGT_E_INVALID, "layout: token channel stride %d < C=%d", tok_ld, C);
  int HW = H * W;
  dim3 grid((HW + 31) / 32, (C + 31) / 32, B * F), block(32, 8);
#define L(TS, TD) ncfhw_tokens_kernel<TS, TD, kToTokens><<<grid, block, 0, st>>>((const TS*)src, (const TS*)add, (TD*)dst, B, C, F, HW, tok_ld)
  if (sdt == MMGT_F32 && ddt == MMGT_F32) L(float, float);
  else if (sdt == MMGT_F32 && ddt == MMGT_BF16) L(float, bf16);
  else if (sdt == MMGT_BF16 && ddt == MMGT_F32) L(bf16, float);
  else if (sdt == MMGT_BF16 && ddt == MMGT_BF16) L(bf16, bf16);
  else { mmgt_set_error("layout: bad dtype"); return MMGT_E_INVALID; }
#undef L
  MMGT_LAUNCH_OK(ctx);
  return 0;
}

extern "C" int mmgt_ncfhw_to_tokens(mmgt_ctx* ctx, const void* src, const void* add, void* dst, int B, int C, int F,
                                    int H, int W, int sdt, int ddt, void* stream) {
  return launch_layout<true>(ctx, src, add, dst, B, C, F, H, W, sdt, ddt, C, (cudaStream_t)stream);
}
extern "C" int mmgt_tokens_to_ncfhw(mmgt_ctx* ctx, const void* src, void* dst, int B, int C, int F, int H, int W,
                                    int src_ld, int sdt, int ddt, void* stream) {
  return launch_layout<false>(ctx, src, nullptr, dst, B, C, F, H, W, sdt, ddt, src_ld, (cudaStream_t)stream);
}

// ---------------------------------------------------------------- nearest x2 upsample, channels-last
template <typename T, int VEC>
__global__ void upsample2x_kernel(const T* __restrict__ x, T* __restrict__ y, int N, int H, int W, int Cv) {
  // one thread per output vector of VEC channels
  typedef typename std::conditional<sizeof(T) * VEC == 16, uint4, typename std::conditional<sizeof(T) * VEC == 8, uint2, T>::type>::type V;
  size_t total = (size_t)N * 2 * H * 2 * W * Cv;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    int cv = i % Cv;
    size_t r = i / Cv;
    int ow = r % (2 * W);
    r /= 2 * W;
    int oh = r % (2 * H);
    int n = r / (2 * H);
    const V* s = reinterpret_cast<const V*>(x) + (((size_t)n * H + (oh >> 1)) * W + (ow >> 1)) * Cv + cv;
    reinterpret_cast<V*>(y)[i] = *s;
  }
}

extern "C" int mmgt_upsample_nearest2x(mmgt_ctx* ctx, const void* x, void* y, int N, int H, int W, int C, int dtype,
                                       void* stream) {
  MMGT_CHECK_ARG(x && y && N > 0 && H > 0 && W > 0 && C > 0, MMGT_E_INVALID, "upsample: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  int esz = dtype == MMGT_F32 ? 4 : 2;
  int vec = 16 / esz;
  bool v = (C % vec == 0) && aligned16(x) && aligned16(y);
  size_t total = (size_t)N * 4 * H * W * (v ? C / vec : C);
  int blocks = (int)std::min<size_t>((total + 255) / 256, (size_t)ctx->num_sms * 16);
  if (dtype == MMGT_F32) {
    if (v) upsample2x_kernel<float, 4><<<blocks, 256, 0, st>>>((const float*)x, (float*)y, N, H, W, C / 4);
    else upsample2x_kernel<float, 1><<<blocks, 256, 0, st>>>((const float*)x, (float*)y, N, H, W, C);
  } else if (dtype == MMGT_BF16) {
    if (v) upsample2x_kernel<bf16, 8><<<blocks, 256, 0, st>>>((const bf16*)x, (bf16*)y, N, H, W, C / 8);
    else upsample2x_kernel<bf16, 1><<<blocks, 256, 0, st>>>((const bf16*)x, (bf16*)y, N, H, W, C);
  } else { mmgt_set_error("upsample: bad dtype"); return MMGT_E_INVALID; }
  MMGT_LAUNCH_OK(ctx);
  return 0;
}

// ---------------------------------------------------------------- im2col 3x3 pad 1 (stride / upsample)
// col[(n,oh,ow), (r*3+s)*C + c] = xin[n, oh*stride+r-1, ow*stride+s-1, c] with xin = (upsampled) x
template <typename T, int VEC>
__global__ void im2col3x3_kernel(const T* __restrict__ x, T* __restrict__ col, int N, int H, int W, int Cv, int stride,
                                 int up) {
  typedef typename std::conditional<sizeof(T) * VEC == 16, uint4, T>::type V;
  const int Hi = up ? 2 * H : H, Wi = up ? 2 * W : W;
  const int Ho = (Hi + 2 - 3) / stride + 1, Wo = (Wi + 2 - 3) / stride + 1;
  size_t total = (size_t)N * Ho * Wo * 9 * Cv;
  V zero;
  memset(&zero, 0, sizeof(V));
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    int cv = i % Cv;
    size_t r = i / Cv;
    int tap = r % 9;
    r /= 9;
    int ow = r % Wo;
    r /= Wo;
    int oh = r % Ho;
    int n = r / Ho;
    int ih = oh * stride + tap / 3 - 1, iw = ow * stride + tap % 3 - 1;
    V v = zero;
    if (ih >= 0 && ih < Hi && iw >= 0 && iw < Wi) {
      if (up) { ih >>= 1; iw >>= 1; }
      v = reinterpret_cast<const V*>(x)[(((size_t)n * H + ih) * W + iw) * Cv + cv];
    }
    reinterpret_cast<V*>(col)[i] = v;
  }
}

extern "C" int mmgt_im2col3x3(mmgt_ctx* ctx, const void* x, void* col, int N, int H, int W, int C, int stride,
                              int upsample2x, int dtype, void* stream) {
  MMGT_CHECK_ARG(x && col && N > 0 && H > 0 && W > 0 && C > 0 && (stride == 1 || stride == 2), MMGT_E_INVALID,
                 "im2col: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  int esz = dtype == MMGT_F32 ? 4 : 2;
  int vec = 16 / esz;
  bool v = (C % vec == 0) && aligned16(x) && aligned16(col);
  const int Hi = upsample2x ? 2 * H : H, Wi = upsample2x ? 2 * W : W;
  const int Ho = (Hi - 1) / stride + 1, Wo = (Wi - 1) / stride + 1;
  size_t total = (size_t)N * Ho * Wo * 9 * (v ? C / vec : C);
  int blocks = (int)std::min<size_t>((total + 255) / 256, (size_t)ctx->num_sms * 16);
  if (dtype == MMGT_F32) {
    if (v) im2col3x3_kernel<float, 4><<<blocks, 256, 0, st>>>((const float*)x, (float*)col, N, H, W, C / 4, stride, upsample2x);
    else im2col3x3_kernel<float, 1><<<blocks, 256, 0, st>>>((const float*)x, (float*)col, N, H, W, C, stride, upsample2x);
  } else if (dtype == MMGT_BF16) {
    if (v) im2col3x3_kernel<bf16, 8><<<blocks, 256, 0, st>>>((const bf16*)x, (bf16*)col, N, H, W, C / 8, stride, upsample2x);
    else im2col3x3_kernel<bf16, 1><<<blocks, 256, 0, st>>>((const bf16*)x, (bf16*)col, N, H, W, C, stride, upsample2x);
  } else { mmgt_set_error("im2col: bad dtype"); return MMGT_E_INVALID; }
  MMGT_LAUNCH_OK(ctx);
  return 0;
}

// ---------------------------------------------------------------- row gather (context-window assembly)
// dst[i, :] = src[idx[i], :] with rows of row_vec 16-byte vectors
__global__ void gather_rows_kernel(const uint4* __restrict__ src, const int32_t* __restrict__ idx, uint4* __restrict__ dst,
                                   int n_out, int64_t row_vec) {
  size_t total = (size_t)n_out * row_vec;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    size_t r = i / row_vec, c = i - r * row_vec;
    dst[i] = src[(size_t)idx[r] * row_vec + c];
  }
}
extern "C" int mmgt_gather_rows(mmgt_ctx* ctx, const void* src, const int32_t* idx, void* dst, int n_out, int64_t row_bytes,
                                void* stream) {
  MMGT_CHECK_ARG(src && idx && dst && n_out > 0 && row_bytes > 0, MMGT_E_INVALID, "gather_rows: bad args");
  MMGT_CHECK_ARG(row_bytes % 16 == 0 && aligned16(src) && aligned16(dst), MMGT_E_ALIGN, "gather_rows: rows must be 16B multiples");
  size_t total = (size_t)n_out * (row_bytes / 16);
  int blocks = (int)std::min<size_t>((total + 255) / 256, (size_t)ctx->num_sms * 16);
  gather_rows_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const uint4*)src, idx, (uint4*)dst, n_out, row_bytes / 16);
  MMGT_LAUNCH_OK(ctx);
  return 0;
}

// ---------------------------------------------------------------- channel padding (conv_in: 4 latent channels -> 64)
// dst[r, :C] = src[r, :], dst[r, C:Cpad] = 0: gives the first convolution a K the tensor-core implicit GEMM takes.
template <typename T>
__global__ void pad_channels_kernel(const T* __restrict__ src, T* __restrict__ dst, int64_t rows, int C, int Cpad) {
  const size_t total = (size_t)rows * Cpad;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = i / Cpad;
    const int c = (int)(i - r * Cpad);
    dst[i] = c < C ? src[r * C + c] : from_f32<T>(0.f);
  }
}
extern "C" int mmgt_pad_channels(mmgt_ctx* ctx, const void* src, void* dst, int64_t rows, int C, int Cpad, int dtype,
                                 void* stream) {
  MMGT_CHECK_ARG(ctx && src && dst && rows > 0 && C > 0 && Cpad >= C, MMGT_E_INVALID, "pad_channels: bad args");
  const size_t total = (size_t)rows * Cpad;
  const int blocks = (int)std::min<size_t>((total + 255) / 256, (size_t)ctx->num_sms * 16);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == MMGT_F32) pad_channels_kernel<float><<<blocks, 256, 0, st>>>((const float*)src, (float*)dst, rows, C, Cpad);
  else if (dtype == MMGT_BF16) pad_channels_kernel<bf16><<<blocks, 256, 0, st>>>((const bf16*)src, (bf16*)dst, rows, C, Cpad);
  else { mmgt_set_error("pad_channels: bad dtype %d", dtype); return MMGT_E_INVALID; }
  MMGT_LAUNCH_OK(ctx);
  return 0;
}

// ---------------------------------------------------------------- small float32 pieces
__global__ void silu_f32_kernel(const float* __restrict__ x, float* __restrict__ y, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float v = x[i];
    y[i] = v / (1.0f + expf(-v));
  }
}
extern "C" int mmgt_silu_f32(mmgt_ctx* ctx, const float* x, float* y, int64_t n, void* stream) {
  MMGT_CHECK_ARG(x && y && n > 0, MMGT_E_INVALID, "silu: bad args");
  int blocks = (int)std::min<int64_t>((n + 255) / 256, 1024);
  silu_f32_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(x, y, (size_t)n);
  MMGT_LAUNCH_OK(ctx);
  return 0;
}

// diffusers get_timestep_embedding: freq_i = exp(-ln(1e4) * i / (half - shift)); [sin | cos], flipped if asked
__global__ void timestep_embedding_kernel(const float* __restrict__ t, float* __restrict__ out, int B, int dim,
                                          int flip, float shift) {
  int half = dim / 2;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * half) return;
  int b = i / half, j = i % half;
  float freq = expf(-9.210340371976184f * (float)j / ((float)half - shift));
  float arg = t[b] * freq;
  float s = sinf(arg), c = cosf(arg);
  float* o = out + (size_t)b * dim;
  if (flip) { o[j] = c; o[half + j] = s; }
  else { o[j] = s; o[half + j] = c; }
}
extern "C" int mmgt_timestep_embedding(mmgt_ctx* ctx, const float* t, float* out, int B, int dim, int flip,
                                       float freq_shift, void* stream) {
  MMGT_CHECK_ARG(t && out && B > 0 && dim > 0 && dim % 2 == 0, MMGT_E_INVALID, "timestep_embedding: bad args");
  int n = B * dim / 2;
  timestep_embedding_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(t, out, B, dim, flip, freq_shift);
  MMGT_LAUNCH_OK(ctx);
  return 0;
}

// ---------------------------------------------------------------- denoise-loop pieces
template <typename T>
__global__ void window_accumulate_kernel(float* __restrict__ acc, const T* __restrict__ pred,
                                         const int32_t* __restrict__ frames, int Bp, int b0, int C, int L, int Fw,
                                         int HW) {
  size_t total = (size_t)Bp * C * Fw * HW;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    int p = i % HW;
    size_t r = i / HW;
    int j = r % Fw;
    r /= Fw;
    int c = r % C;
    int b = r / C;
    size_t o = (((size_t)(b0 + b) * C + c) * L + frames[j]) * HW + p;
    acc[o] += to_f32(pred[i]);
  }
}
extern "C" int mmgt_window_accumulate(mmgt_ctx* ctx, float* acc, const void* pred, const int32_t* frames, int Bp,
                                      int b0, int C, int L, int Fw, int HW, int pdt, void* stream) {
  MMGT_CHECK_ARG(acc && pred && frames && Bp > 0 && C > 0 && L > 0 && Fw > 0 && HW > 0, MMGT_E_INVALID,
                 "window_accumulate: bad args");
  size_t total = (size_t)Bp * C * Fw * HW;
  int blocks = (int)std::min<size_t>((total + 255) / 256, (size_t)ctx->num_sms * 8);
  if (pdt == MMGT_F32) window_accumulate_kernel<float><<<blocks, 256, 0, (cudaStream_t)stream>>>(acc, (const float*)pred, frames, Bp, b0, C, L, Fw, HW);
  else if (pdt == MMGT_BF16) window_accumulate_kernel<bf16><<<blocks, 256, 0, (cudaStream_t)stream>>>(acc, (const bf16*)pred, frames, Bp, b0, C, L, Fw, HW);
  else { mmgt_set_error("window_accumulate: bad dtype"); return MMGT_E_INVALID; }
  MMGT_LAUNCH_OK(ctx);
  return 0;
}

__global__ void cfg_ddim_kernel(float* __restrict__ lat, const float* __restrict__ acc, const float* __restrict__ inv_count,
                                int C, int L, int HW, int cfg, float g, float cx, float cv) {
  size_t per = (size_t)C * L * HW;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < per; i += (size_t)gridDim.x * blockDim.x) {
    int f = (i / HW) % L;
    float ic = inv_count[f];
    float u = acc[i] * ic;
    float v = u;
    if (cfg) {
      float c = acc[per + i] * ic;
      v = u + g * (c - u);
    }
    lat[i] = cx * lat[i] + cv * v;
  }
}
extern "C" int mmgt_cfg_ddim_step(mmgt_ctx* ctx, float* latents, const float* noise_acc, const float* inv_count, int C,
                                  int L, int HW, int cfg, float guidance, float cx, float cv, void* stream) {
  MMGT_CHECK_ARG(latents && noise_acc && inv_count && C > 0 && L > 0 && HW > 0, MMGT_E_INVALID, "cfg_ddim: bad args");
  size_t per = (size_t)C * L * HW;
  int blocks = (int)std::min<size_t>((per + 255) / 256, (size_t)ctx->num_sms * 8);
  cfg_ddim_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(latents, noise_acc, inv_count, C, L, HW, cfg, guidance, cx, cv);
  MMGT_LAUNCH_OK(ctx);
  return 0;
}
