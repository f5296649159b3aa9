// Shared helpers for the mmgt_b200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <unordered_set>

#include "../../include/mmgt_b200.h"

struct mmgt_ctx {
  int device;
  int num_sms;
  int max_smem_optin;
  int use_tc;                 // tcgen05 kernels enabled for bf16
  int use_bres;               // weight-stationary GEMM variant enabled (small K)
  int use_pdl;                // launch with programmatic stream serialization (kernel prologues overlap the previous tail)
  long long launches;         // kernels launched through this context
  void* encode_tiled;         // PFN of cuTensorMapEncodeTiled (resolved lazily through the runtime)
  int strict_tc;              // bf16 requests that no tensor-core kernel covers fail (MMGT_E_UNSUPPORTED) instead of
                              // running on the CUDA-core kernels
  int residual_mma;           // residual epilogues of the streaming GEMM / conv kernels through [R | I] k-blocks (default 1; A/B)
  void* identity;             // 256 x 256 bf16 identity on this device (B operand of those k-blocks), owned by the context
  int tma_store;              // lean epilogues write their tiles through TMA stores (default 1; 0 = per-lane stores, A/B)
  int lean_epilogue;          // tensor-core GEMM / conv: specialised straight-line epilogues where the options allow (default 1; A/B)
  int temporal_rows;          // temporal attention (head dim <= 80) on the row-coalesced cp.async kernel (default 1; A/B)
  int ln_persist;             // LayerNorm on one exact wave of persistent blocks with next-row prefetch (A/B)
  int attn_q256;              // head dim <= 64: 256 queries per CTA, one query tile + one MMA-issuing warp per softmax group (A/B)
  int attn_packed;            // attention softmax loops on packed fp32 pairs (FFMA2 / FADD2); same arithmetic (A/B)
  int attn_persist;           // head dim <= 64: one persistent CTA per SM walking (frame, head, query tile) items: 0 off (default: measured faster), 1 when
                              // there are more items than SMs, n >= 2 always and on at most n CTAs (tests; A/B)
  int attn_v2;                // head dim <= 64: 1 = the three-S-buffer attention kernel, 0 = the two-buffer kernel (default; A/B)
  int conv_implicit_all;      // stride-2 / upsampling convolutions as implicit GEMMs too (default); 0 = im2col staging (A/B)
  int gn_split;               // GroupNorm as statistics kernel + normalise kernel (default) instead of the fused spin-barrier kernel
  int geglu_exact;            // tensor-core GEGLU epilogue uses the erf (Abramowitz-Stegun) form instead of the logistic fit
  long long simt_launches;    // bf16 operator calls that DID take a CUDA-core kernel while tensor cores were enabled
  std::unordered_set<const void*> smem_optin_done;   // kernels whose dynamic-smem limit was raised on THIS device
};

// cudaFuncAttributeMaxDynamicSharedMemorySize is per device: remember it per context, not per process.
template <typename K>
inline cudaError_t mmgt_smem_optin(mmgt_ctx* ctx, K* kernel, int bytes) {
  const void* key = reinterpret_cast<const void*>(kernel);
  if (ctx->smem_optin_done.count(key)) return cudaSuccess;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) ctx->smem_optin_done.insert(key);
  return e;
}

// A bf16 request is about to run on a CUDA-core kernel although tensor cores are enabled: count it, refuse it in strict mode.
#define MMGT_SIMT_FALLBACK(ctx, what)                                                                         \
  do {                                                                                                        \
    if ((ctx)->use_tc) {                                                                                      \
      if ((ctx)->strict_tc) {                                                                                 \
        mmgt_set_error("%s: no tensor-core kernel covers this bf16 request (strict mode, mmgt_ctx_flag 4)", what); \
        return MMGT_E_UNSUPPORTED;                                                                            \
      }                                                                                                       \
      (ctx)->simt_launches++;                                                                                 \
    }                                                                                                         \
  } while (0)

void mmgt_set_error(const char* fmt, ...);

#define MMGT_CHECK_ARG(cond, code, ...)      \
  do {                                       \
    if (!(cond)) {                           \
      mmgt_set_error(__VA_ARGS__);           \
      return (code);                         \
    }                                        \
  } while (0)

#define MMGT_CUDA_OK(expr)                                                          \
  do {                                                                              \
    cudaError_t _e = (expr);                                                        \
    if (_e != cudaSuccess) {                                                        \
      mmgt_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return (int)_e;                                                               \
    }                                                                               \
  } while (0)

// after a kernel launch: count it and surface launch errors (never synchronises)
#define MMGT_LAUNCH_OK(ctx)                                                         \
  do {                                                                              \
    (ctx)->launches++;                                                              \
    cudaError_t _e = cudaGetLastError();                                            \
    if (_e != cudaSuccess) {                                                        \
      mmgt_set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
      return (int)_e;                                                               \
    }                                                                               \
  } while (0)

typedef __nv_bfloat16 bf16;

// ---- Programmatic dependent launch (PDL).  A step is ~23 000 short dependent kernels replayed from one CUDA graph;
// with the attribute below kernel i+1 is scheduled while kernel i drains: its CTAs become resident as soon as
// resources free up, run their prologue (barrier init, TMEM allocation, tensor-map fetch) and block in
// griddepcontrol.wait until every CTA of kernel i has finished and its stores are visible.  Every kernel launched
// through mmgt_launch() calls pdl_prologue() before its first global-memory access, which keeps plain stream
// semantics (the attribute without the wait would be a race).  Launched without the attribute (flag 3 off, or by
// a plain <<<>>>) both instructions are no-ops.
__device__ __forceinline__ void pdl_prologue() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}

// One lane of a converged warp.  The TMA-producer and MMA-issuer roles run under `warp == k && elect_one_sync()` with a
// warp-uniform `warp`: ptxas then KNOWS a single lane is active and feeds tcgen05.mma / cp.async.bulk their uniform-register
// operands directly.  Under `threadIdx.x == 32` it wrapped every UTCHMMA in an ELECT / R2UR.BROADCAST / BRA.U.ANY
// "waterfall" loop with the descriptor arithmetic serialised in front of it (~17 dependent instructions, 70-100 clocks per
// MMA): the issuing thread, not the tensor pipe, paced the attention kernel (11 MMAs per 128 x 128 tile) and the narrow-tile
// GEMMs (profiles/r2_mma_issue.md).
__device__ __forceinline__ uint32_t elect_one_sync() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(pred));
  return pred;
}
__device__ __forceinline__ int uniform_warp_index() { return __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0); }

template <typename... KArgs, typename... Args>
inline cudaError_t mmgt_launch(const mmgt_ctx* ctx, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                               cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = ctx->use_pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f32<bf16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + __expf(-x)); }
// exact (erf) GELU as F.gelu default
__device__ __forceinline__ float gelu_erf_f(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Row exchange (include/mmgt_b200.h, mmgt_row_exchange): destination of source row m, as an element pointer.
template <typename T>
__device__ __forceinline__ T* exchange_row_ptr(const mmgt_row_exchange& ex, int m) {
  const int Fl = ex.F / ex.k, Tc = ex.T / ex.k;
  int s;
  int64_t row;
  if (ex.direction == 1) {
    const int n_loc = m / ex.T, t = m - n_loc * ex.T;
    const int b = n_loc / Fl, f_loc = n_loc - b * Fl;
    s = t / Tc;
    row = (int64_t)(b * ex.F + ex.my * Fl + f_loc) * Tc + (t - s * Tc);
  } else {
    const int n = m / Tc, t_loc = m - n * Tc;
    const int b = n / ex.F, f = n - b * ex.F;
    s = f / Fl;
    row = (int64_t)(b * Fl + (f - s * Fl)) * ex.T + ex.my * Tc + t_loc;
  }
  return reinterpret_cast<T*>(ex.peer_base[s]) + row * ex.ld;
}
int mmgt_row_exchange_check(const mmgt_row_exchange* ex, int64_t rows, const char* who);

static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
static inline bool aligned32(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 31u) == 0; }

// dtype dispatch: calls fn(T{}) with T = float or bf16
#define MMGT_DISPATCH_DTYPE(dtype, T, ...)                          \
  do {                                                              \
    if ((dtype) == MMGT_F32) { using T = float; __VA_ARGS__; }      \
    else if ((dtype) == MMGT_BF16) { using T = bf16; __VA_ARGS__; } \
    else { mmgt_set_error("bad dtype %d", (int)(dtype)); return MMGT_E_INVALID; } \
  } while (0)

// tensor-core paths (gemm_tc.cu)
int mmgt_gemm_tc(mmgt_ctx* ctx, const mmgt_gemm_params* p, cudaStream_t st);
bool mmgt_gemm_tc_supported(const mmgt_ctx* ctx, const mmgt_gemm_params* p);
int mmgt_conv3x3_tc(mmgt_ctx* ctx, const mmgt_conv3x3_params* p, cudaStream_t st);
bool mmgt_conv3x3_tc_supported(const mmgt_ctx* ctx, const mmgt_conv3x3_params* p);
// attention_tc.cu
int mmgt_attention_tc(mmgt_ctx* ctx, const mmgt_attention_params* p, cudaStream_t st);
bool mmgt_attention_tc_supported(const mmgt_ctx* ctx, const mmgt_attention_params* p);
