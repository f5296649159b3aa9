// Motion-mask pyramid: bit-exact Pillow 8-bit bilinear resize (ImagingResample: horizontal pass, then vertical
// pass, 22-bit fixed-point coefficients, uint8 intermediate) followed by ToTensor (u8 / 255 in float32).
// Coefficients are computed on the host in double exactly as Pillow's precompute_coeffs does and travel as
// kernel parameters, so the device only performs integer arithmetic.
#include <math.h>

#include "common.cuh"

namespace {

constexpr int PRECISION_BITS = 32 - 8 - 2;
constexpr int MAX_OUT = 128;
constexpr int MAX_COEF = 640;

struct ResizeTable {
  int ksize;
  int bounds[MAX_OUT][2];   // xmin, count
  int kk[MAX_COEF];         // [out][ksize]
};

// triangle ("bilinear") filter, support 1.0
bool build_table(int in_size, int out_size, ResizeTable* t) {
  const double scale = (double)in_size / out_size;
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = 1.0 * filterscale;
  const int ksize = (int)ceil(support) * 2 + 1;
  if (out_size > MAX_OUT || out_size * ksize > MAX_COEF) return false;
  t->ksize = ksize;
  for (int xx = 0; xx < out_size; ++xx) {
    const double center = (xx + 0.5) * scale;
    const double ss = 1.0 / filterscale;
    int xmin = (int)(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    double w[64];
    double ww = 0.0;
    for (int x = 0; x < xmax; ++x) {
      double v = (x + xmin - center + 0.5) * ss;
      if (v < 0) v = -v;
      double wv = v < 1.0 ? 1.0 - v : 0.0;
      w[x] = wv;
      ww += wv;
    }
    for (int x = 0; x < xmax; ++x)
      if (ww != 0.0) w[x] /= ww;
    for (int x = 0; x < ksize; ++x) {
      double c = x < xmax ? w[x] * (double)(1 << PRECISION_BITS) : 0.0;
      t->kk[xx * ksize + x] = c < 0 ? (int)(c - 0.5) : (int)(c + 0.5);
    }
    t->bounds[xx][0] = xmin;
    t->bounds[xx][1] = xmax;
  }
  return true;
}

__device__ __forceinline__ uint8_t clip8(int v) {
  v >>= PRECISION_BITS;
  return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// horizontal: src (L*H, Win) -> dst (L*H, Wout)
__global__ void resize_h_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int rows, int Win, int Wout,
                                const __grid_constant__ ResizeTable t) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * Wout) return;
  int r = i / Wout, xx = i % Wout;
  int xmin = t.bounds[xx][0], cnt = t.bounds[xx][1];
  int acc = 1 << (PRECISION_BITS - 1);
  for (int x = 0; x < cnt; ++x) acc += (int)src[(size_t)r * Win + xmin + x] * t.kk[xx * t.ksize + x];
  dst[i] = clip8(acc);
}

// vertical: src (L, Hin, W) -> dst (L, Hout, W); also emits float = offset + u8/255
__global__ void resize_v_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst_u8, float* __restrict__ dst_f,
                                int L, int Hin, int Hout, int W, float offset, int identity,
                                const __grid_constant__ ResizeTable t) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= L * Hout * W) return;
  int x = i % W, yy = (i / W) % Hout, l = i / (W * Hout);
  uint8_t v;
  if (identity) {
    v = src[i];
  } else {
    int ymin = t.bounds[yy][0], cnt = t.bounds[yy][1];
    int acc = 1 << (PRECISION_BITS - 1);
    for (int y = 0; y < cnt; ++y) acc += (int)src[((size_t)l * Hin + ymin + y) * W + x] * t.kk[yy * t.ksize + y];
    v = clip8(acc);
  }
  if (dst_u8) dst_u8[i] = v;
  if (dst_f) dst_f[i] = offset + (float)v / 255.0f;
}

}  // namespace

extern "C" int mmgt_mask_resize(mmgt_ctx* ctx, const uint8_t* src, uint8_t* tmp, uint8_t* out_u8, float* out_f32, int L,
                                int Hs, int Ws, int S, float offset, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  MMGT_CHECK_ARG(ctx && src && tmp && (out_u8 || out_f32) && L > 0 && Hs > 0 && Ws > 0 && S > 0, MMGT_E_INVALID,
                 "mask_resize: bad args");
  static thread_local ResizeTable th, tv;
  const uint8_t* vsrc = src;
  if (S != Ws) {
    MMGT_CHECK_ARG(build_table(Ws, S, &th), MMGT_E_UNSUPPORTED, "mask_resize: %d -> %d exceeds the coefficient table", Ws, S);
    int n = L * Hs * S;
    resize_h_kernel<<<(n + 255) / 256, 256, 0, st>>>(src, tmp, L * Hs, Ws, S, th);
    MMGT_LAUNCH_OK(ctx);
    vsrc = tmp;
  }
  const int identity = (S == Hs);
  if (!identity) MMGT_CHECK_ARG(build_table(Hs, S, &tv), MMGT_E_UNSUPPORTED, "mask_resize: %d -> %d exceeds the coefficient table", Hs, S);
  int n = L * S * S;
  resize_v_kernel<<<(n + 255) / 256, 256, 0, st>>>(vsrc, out_u8, out_f32, L, Hs, S, S, offset, identity, tv);
  MMGT_LAUNCH_OK(ctx);
  return 0;
}
