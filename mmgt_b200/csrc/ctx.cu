// Context, error reporting, ABI version.
#include <stdarg.h>
#include <string.h>
#include <vector>

#include "common.cuh"

static thread_local char g_err[512] = "";

void mmgt_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* mmgt_last_error(void) { return g_err; }
extern "C" int mmgt_abi_version(void) { return 3; }

extern "C" int mmgt_ctx_create(mmgt_ctx** out, int device) {
  MMGT_CHECK_ARG(out != nullptr, MMGT_E_INVALID, "mmgt_ctx_create: out is NULL");
  MMGT_CUDA_OK(cudaSetDevice(device));
  cudaDeviceProp prop;
  MMGT_CUDA_OK(cudaGetDeviceProperties(&prop, device));
  MMGT_CHECK_ARG(prop.major == 10, MMGT_E_UNSUPPORTED,
                 "mmgt_b200 is built for sm_100a only; device %d is sm_%d%d", device, prop.major, prop.minor);
  mmgt_ctx* c = new mmgt_ctx();
  c->device = device;
  c->num_sms = prop.multiProcessorCount;
  c->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
  c->use_tc = 1;
  c->use_bres = 1;
  c->use_pdl = 0;
  c->launches = 0;
  c->encode_tiled = nullptr;
  c->strict_tc = 0;
  c->simt_launches = 0;
  c->geglu_exact = 0;
  c->gn_split = 1;
  c->conv_implicit_all = 1;
  c->temporal_rows = 1;
  c->lean_epilogue = 1;
  c->tma_store = 1;
  c->residual_mma = 1;
  c->identity = nullptr;
  {   // the one device allocation the context owns (128 KB): created here so that no launch ever allocates (graph capture)
    std::vector<uint16_t> eye(256 * 256, 0);
    for (int i = 0; i < 256; ++i) eye[i * 256 + i] = 0x3F80;      // bf16 1.0
    if (cudaMalloc(&c->identity, eye.size() * 2) != cudaSuccess ||
        cudaMemcpy(c->identity, eye.data(), eye.size() * 2, cudaMemcpyHostToDevice) != cudaSuccess) {
      cudaGetLastError();
      if (c->identity) cudaFree(c->identity);
      c->identity = nullptr;                                       // residual_mma then stays unused
    }
  }
  c->attn_persist = 0;  // measured (profiles/r2_attention_v2.md): 2134 vs 1815 us per level-0 launch, 452 vs 442 ms per step in favour of one item per CTA
  c->attn_q256 = 5;    // measured (profiles/r2_attention_q256.md): 1476 vs 1815 us per level-0 launch, 413 vs 435 ms per step
  c->attn_packed = 1;
  c->ln_persist = 0;
  c->attn_v2 = 0;     // measured (profiles/r2_ab_flags.md): 475 vs 482 ms per step in favour of the two-buffer kernel
  *out = c;
  return 0;
}

extern "C" int mmgt_ctx_destroy(mmgt_ctx* ctx) {
  if (ctx && ctx->identity) cudaFree(ctx->identity);
  delete ctx;
  return 0;
}

extern "C" int64_t mmgt_ctx_flag(mmgt_ctx* ctx, int flag, int64_t value) {
  if (!ctx) return MMGT_E_INVALID;
  if (flag == 0) {
    if (value >= 0) ctx->use_tc = value ? 1 : 0;
    return ctx->use_tc;
  }
  if (flag == 1) return ctx->launches;
  if (flag == 2) {
    if (value >= 0) ctx->use_bres = value ? 1 : 0;
    return ctx->use_bres;
  }
  if (flag == 3) {
    if (value >= 0) ctx->use_pdl = value ? 1 : 0;
    return ctx->use_pdl;
  }
  if (flag == 4) {
    if (value >= 0) ctx->strict_tc = value ? 1 : 0;
    return ctx->strict_tc;
  }
  if (flag == 5) return ctx->simt_launches;
  if (flag == 17) {
    if (value >= 0) ctx->ln_persist = (int)value;
    return ctx->ln_persist;
  }
  if (flag == 15) {
    if (value >= 0) ctx->attn_q256 = (int)value;
    return ctx->attn_q256;
  }
  if (flag == 16) {
    if (value >= 0) ctx->attn_packed = value ? 1 : 0;
    return ctx->attn_packed;
  }
  if (flag == 14) {
    if (value >= 0) ctx->attn_persist = (int)value;      // 0 off, 1 auto, n >= 2: always, on at most n CTAs (tests)
    return ctx->attn_persist;
  }
  if (flag == 13) {
    if (value >= 0) ctx->residual_mma = value ? 1 : 0;
    return ctx->residual_mma;
  }
  if (flag == 12) {
    if (value >= 0) ctx->tma_store = value ? 1 : 0;
    return ctx->tma_store;
  }
  if (flag == 11) {
    if (value >= 0) ctx->lean_epilogue = value ? 1 : 0;
    return ctx->lean_epilogue;
  }
  if (flag == 10) {
    if (value >= 0) ctx->temporal_rows = value ? 1 : 0;
    return ctx->temporal_rows;
  }
  if (flag == 9) {
    if (value >= 0) ctx->attn_v2 = value ? 1 : 0;
    return ctx->attn_v2;
  }
  if (flag == 8) {
    if (value >= 0) ctx->conv_implicit_all = value ? 1 : 0;
    return ctx->conv_implicit_all;
  }
  if (flag == 7) {
    if (value >= 0) ctx->gn_split = value ? 1 : 0;
    return ctx->gn_split;
  }
  if (flag == 6) {
    if (value >= 0) ctx->geglu_exact = value ? 1 : 0;
    return ctx->geglu_exact;
  }
  return MMGT_E_INVALID;
}
