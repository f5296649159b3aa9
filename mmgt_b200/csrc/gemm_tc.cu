// tcgen05 / TMEM / TMA GEMM and implicit-GEMM 3x3 convolution for sm_100a (bf16 operands, fp32 accumulate).
//
//   D[M, N] = epilogue( A[M, K] * W[N, K]^T )
//
// One persistent CTA per SM, warp-specialised:
//   warp 0 (one lane)  TMA producer: A / W tiles -> 128B-swizzled shared memory, mbarrier complete_tx
//   warp 1 (one lane)  tcgen05.mma issuer: 128 x BN x 16 UMMAs into a double-buffered TMEM accumulator
//   warps 2..9         epilogue: tcgen05.ld -> bias / row-scale / row-bias / residual / GEGLU -> global stores
// Pipelines: smem full/empty ring (TMA <-> MMA) and TMEM full/empty pair (MMA <-> epilogue), so the epilogue
// of tile i overlaps the main loop of tile i+1.
//
// The convolution variant replaces the A loads by 4-D TMA boxes (C, W, H, N) shifted by the filter tap;
// out-of-image coordinates are zero-filled by the TMA unit, which implements padding = 1 for free, and the
// K loop runs over 9 taps x Cin/64 channel blocks against the (Cout, 3, 3, Cin) weight viewed as (Cout, 9*Cin).
#include <cuda.h>

#include "common.cuh"

namespace {

constexpr int BM = 128;
constexpr int BK = 64;             // 64 bf16 = 128 bytes = one swizzle atom row
constexpr int UMMA_K = 16;
constexpr int NUM_THREADS = 320;            // TMA warp + MMA warp + 8 epilogue warps
constexpr int A_STAGE_BYTES = BM * BK * 2;
constexpr int EPI_VEC_BYTES = 2 * 3 * 256 * 4;   // [accumulator stage][bias | colsum | row bias][256 columns] floats
// Shared memory after the ring: barriers (256 B) + epilogue vectors, padded to a 1 KB boundary, then the staging boxes
// of the TMA-store epilogue: per epilogue warp 32 rows x 64 B (two 16-column chunks) or x 32 B (one chunk).
constexpr int EPI_FIXED_BYTES = 7168;
static_assert(256 + EPI_VEC_BYTES <= EPI_FIXED_BYTES, "epilogue vectors overflow their slot");
__host__ __device__ constexpr int epi_out_bytes(int tma_store) { return tma_store == 2 ? 16384 : tma_store == 1 ? 8192 : 0; }

__host__ __device__ constexpr int acc_stride(int bn) { return bn <= 32 ? 32 : bn <= 64 ? 64 : bn <= 128 ? 128 : 256; }
__host__ __device__ constexpr int num_stages(int bn) { return bn >= 256 ? 4 : bn >= 160 ? 5 : bn >= 128 ? 6 : 8; }

struct TcArgs {
  int M, N_out, num_m_tiles, num_n_tiles, num_k_blocks;
  // epilogue
  const float* bias;
  const float* rowscale;
  const float* rowbias;
  const bf16* residual;
  bf16* D;
  int64_t ldd, ldr;
  int rows_per_group;
  float alpha;
  int64_t ld_rowbias;       // elements between rowbias rows
  int rowbias_mod;          // > 0: rowbias row = (m / rows_per_group) % rowbias_mod
  const float* rowstats;    // fused LayerNorm: (M, 2) [mean, rstd]; D = rstd * (acc - mean * colsum[n]) + bias[n]
  const float* colsum;
  int act;                  // 0 none, 1 SiLU, 2 ReLU
  int rb_tile_rows;         // > 0: every M tile lies inside ONE row-bias group of this many TILE-SPACE rows (a multiple of
                            // 128), so the tile's row-bias vector is staged in shared memory with bias / colsum
  // conv geometry (CONV only).  H, W span the TILE space: the output image for conv_mode 0 (3x3, stride 1: same as the
  // input) and 1 (stride 2: A boxes are fetched with a TMA traversal stride of 2 from coordinate 2*o + tap - 1); the
  // INPUT image for conv_mode 2 (nearest x2 upsample + 3x3 = four 2x2-tap convolutions on the low-resolution input with
  // pre-summed weights, one per output parity (a, b); tile index = 4 * input tile + parity, K = 4 taps x Cin, output row
  // (n, 2h + a, 2w + b)).
  int H, W, cin_blocks;
  int conv_mode;
  // CONV, patch tiles: when no run of 128 consecutive (n, h, w) rows is a TMA box (W = 96, 48, 24, 12: the 768^2 / 384^2
  // latents of config 5) an M tile is a (pf frames) x (ph rows) x (pw columns) patch, pw * ph * pf = 128.  pw == 0:
  // tiles are 128 consecutive rows.
  int pw, ph, pf, n_frames;
  // BRES only: A ring depth (runtime: what is left of shared memory after the resident weight tile)
  int stages;
  int wide_io;   // D / residual rows are 32-byte aligned: 256-bit epilogue loads and stores
  int epi;       // epilogue code variant picked by the host (see the EPI template parameter)
  // Lean epilogues only: output through TMA.  The row-per-lane 32-byte stores cost ~80-110 clk of LSU / L1 time per warp
  // instruction (32 lanes = 32 different lines) and paced every GEMM of the path (profiles/r2_k_sweep_store_ablation.txt:
  // 960x320 74.6 us with, 49.2 us without its stores).  Instead each warp writes its 32-row x 16-column chunks (bf16) into
  // a swizzled shared-memory box and one lane hands the box to cp.async.bulk.tensor (tmD2: two chunks, 64-byte rows; tmD1:
  // one chunk, the tail of an odd chunk count).  0: off.  1: one-chunk boxes only (shared memory is short).  2: both.
  // Needs tiles of 128 consecutive output rows (plain GEMM, box-tiled conv); rows >= M / columns >= N are clipped by TMA.
  int tma_store;
  // Residual through the tensor cores (streaming kernels): after the K blocks of A x W^T the producer feeds res_kblocks
  // more k-blocks whose A operand is the residual tile itself (rows x 64 columns, by TMA from the residual matrix: the
  // kernel parameter tmD2 is that map) and whose B operand is a slice of an identity matrix (tmD1), so the accumulator
  // ends up holding A W^T + R exactly (bf16 x 1.0 into fp32) and the epilogue neither loads nor adds anything per lane:
  // the row-per-lane residual loads cost the LSU as much as the stores (profiles/r2_k_sweep_store_ablation.txt).
  int res_kblocks;
  // Row exchange (multi-GPU frame-shard <-> token-shard switch around the motion modules): when ex.direction != 0
  // the epilogue stores row m into the receive buffer of the shard that owns it -- local or a peer's, over NVLink --
  // so the all-to-all is part of the GEMM that produces the rows and overlaps its main loop tile by tile.
  mmgt_row_exchange ex;
};

// ------------------------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done;
  do {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(smem_src), "r"(c0),
               "r"(c1)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void sts_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, 128B swizzle: rows are 128 bytes, 8-row groups 1024 bytes apart (SBO), LBO unused (1), version 1.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// kind::f16 instruction descriptor: D=f32, A=B=bf16, both K-major, M=128, N=bn
__host__ __device__ constexpr uint32_t make_idesc(int bn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld_x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
// Scheduling fence on 16 registers: everything that consumes r[] is ordered after this point (and r[] stays live across
// it), so a tcgen05.ld issued just before it into the OTHER register set really is in flight while r[] is processed --
// without it ptxas sinks the next load below the arithmetic and reuses the same registers (no overlap at all).
__device__ __forceinline__ void pin16(uint32_t (&r)[16]) {
  asm volatile("" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
               "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void store16_bf16(bf16* dst, const float (&v)[16]) {
  uint4 o[2];
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(o);
#pragma unroll
  for (int i = 0; i < 8; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  reinterpret_cast<uint4*>(dst)[0] = o[0];
  reinterpret_cast<uint4*>(dst)[1] = o[1];
}
__device__ __forceinline__ void load16_bf16(const bf16* src, float (&v)[16]) {
  uint4 o[2];
  o[0] = reinterpret_cast<const uint4*>(src)[0];
  o[1] = reinterpret_cast<const uint4*>(src)[1];
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(o);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float2 f = __bfloat1622float2(h[i]);
    v[2 * i] = f.x; v[2 * i + 1] = f.y;
  }
}


// 32-byte (16 x bf16) global accesses.  sm_100 has 256-bit LDG / STG: one instruction (and one L1 tag look-up per
// lane) instead of two -- the epilogue is row-per-thread, so every lane of a warp touches a different line and the
// LSU, not the tensor pipe, paces small-K GEMMs.  `wide` needs 32-byte aligned addresses.
struct Row32 { uint32_t w[8]; };
__device__ __forceinline__ Row32 ld_row32(const bf16* p, bool wide) {
  Row32 r;
  if (wide) {
    asm volatile("ld.global.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]), "=r"(r.w[4]), "=r"(r.w[5]), "=r"(r.w[6]), "=r"(r.w[7])
                 : "l"(p));
  } else {
    const uint4 a = reinterpret_cast<const uint4*>(p)[0], b = reinterpret_cast<const uint4*>(p)[1];
    r.w[0] = a.x; r.w[1] = a.y; r.w[2] = a.z; r.w[3] = a.w; r.w[4] = b.x; r.w[5] = b.y; r.w[6] = b.z; r.w[7] = b.w;
  }
  return r;
}
__device__ __forceinline__ void st_row32(bf16* p, const float (&v)[16], bool wide) {
  uint32_t w[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    w[i] = *reinterpret_cast<uint32_t*>(&h);
  }
  if (wide) {
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(w[0]), "r"(w[1]), "r"(w[2]),
                 "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7])
                 : "memory");
  } else {
    reinterpret_cast<uint4*>(p)[0] = make_uint4(w[0], w[1], w[2], w[3]);
    reinterpret_cast<uint4*>(p)[1] = make_uint4(w[4], w[5], w[6], w[7]);
  }
}

// Predicated forms for the straight-line epilogue: no branch around the instruction, so ptxas keeps the whole tile in one
// basic block and overlaps neighbouring chunks.
__device__ __forceinline__ void st_row32_if(bool pred, bf16* p, const float (&v)[16]) {
  uint32_t w[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    w[i] = *reinterpret_cast<uint32_t*>(&h);
  }
  asm volatile(
      "{\n\t"
      ".reg .pred q;\n\t"
      "setp.ne.b32 q, %9, 0;\n\t"
      "@q st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n\t"
      "}" ::"l"(p),
      "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]), "r"((int)pred)
      : "memory");
}
__device__ __forceinline__ void ld_row32_if(bool pred, Row32& r, const bf16* p) {
  asm volatile(
      "{\n\t"
      ".reg .pred q;\n\t"
      "setp.ne.b32 q, %9, 0;\n\t"
      "@q ld.global.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n\t"
      "}"
      : "+r"(r.w[0]), "+r"(r.w[1]), "+r"(r.w[2]), "+r"(r.w[3]), "+r"(r.w[4]), "+r"(r.w[5]), "+r"(r.w[6]), "+r"(r.w[7])
      : "l"(p), "r"((int)pred));
}

__device__ __forceinline__ void lds16_f32(const float* src, float (&v)[16]) {    // shared memory, warp-uniform address (broadcast)
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 f = reinterpret_cast<const float4*>(src)[i];
    v[4 * i] = f.x; v[4 * i + 1] = f.y; v[4 * i + 2] = f.z; v[4 * i + 3] = f.w;
  }
}
__device__ __forceinline__ void load16_f32(const float* src, float (&v)[16]) {   // 64-byte aligned, warp-uniform address
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 f = __ldg(reinterpret_cast<const float4*>(src) + i);
    v[4 * i] = f.x; v[4 * i + 1] = f.y; v[4 * i + 2] = f.z; v[4 * i + 3] = f.w;
  }
}

// erf-GELU for the tensor-core GEGLU epilogue: 0.5 x (1 + erf(x / sqrt 2)) with erfc(|z|) from Abramowitz-Stegun
// 7.1.26 (|erf error| <= 1.5e-7, far below the bf16 output rounding) -- 2 MUFU + ~14 FP32 ops instead of the
// ~35-instruction branchy erff(); the fp32 kernels keep erff().
__device__ __forceinline__ float gelu_erf_fast(float x) {
  const float z = fabsf(x) * 0.70710678118654752f;
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.f)));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * x * -0.72134752044448170f));   // exp(-x^2 / 2)
  const float erfc_abs = p * t * e;                       // erfc(|x| / sqrt 2)
  const float cdf2 = x < 0.f ? erfc_abs : 2.f - erfc_abs; // 1 + erf(x / sqrt 2)
  return 0.5f * x * cdf2;
}

// GEGLU epilogue of the bf16 tier: x * Phi(x) with Phi(x) = 1 / (1 + exp(-x (c1 + c3 z + c5 z^2))), z = min(x^2, 42):
// an odd polynomial fit of logit(Phi) (least squares on the relative error of x Phi(x) over |x| <= 12).  |error| <=
// 6.3e-5 absolute and <= 8.2e-4 relative (0.4 bf16 half-ulps) against the erf form, correct tails (Phi -> 0 / 1 with
// relative accuracy), 2 MUFU + 8 FP32 operations per element instead of 2 + ~17: at C = 320 the GEGLU GEMM is paced by
// the issue slots of its epilogue (128 x 128 outputs per tile against 5 k-blocks of MMA), not by the tensor pipe.
// Coefficients are pre-multiplied by -log2(e).  gelu_erf_fast() above stays for reference / A-B (MMGT_GEGLU_EXACT).
__device__ __forceinline__ float gelu_logistic(float x) {
  const float z = fminf(x * x, 42.f);
  float p = fmaf(0.0011339103803038597f, z, -0.10764378309249878f);
  p = fmaf(p, z, -2.300089120864868f);
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(p * x));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.f));
  return x * r;
}
__device__ __forceinline__ float act_apply(float v, int act) {
  if (act == 1) {
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(v * -1.4426950408889634f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.f));
    return v * r;
  }
  if (act == 2) return fmaxf(v, 0.f);
  return v;
}

// ------------------------------------------------------------------------------------------- kernel
// BRES ("weight-stationary"): for small K the whole (BN x K) weight tile of this CTA stays resident in shared
// memory, every CTA keeps one n-block for its lifetime and only A tiles stream through the ring.  The
// L2 -> SM fabric (~45 B/clk/SM), not the tensor pipe, bounds a 128 x BN tile that re-reads B per tile:
// (128 + BN) * 128 B per k-block vs. 128 * 128 B with B resident.
// EPI selects the epilogue code: 0 = every option behind (warp-uniform) run-time branches; 1 = "lean": bias (+ the
// tile's staged row bias, pre-added in shared memory) [+ GEGLU], 256-bit stores; 2 = lean + residual; 3 / 4 = 1 / 2 with
// the output tiles going through TMA stores (TcArgs::tma_store; only 4 is instantiated, see mmgt_gemm_tc).  The lean forms
// are one straight-line block per tile: the general form's per-chunk option branches kept ptxas from interleaving
// chunks and cost the epilogue warps 24 % "no instruction" + 13 % branch-resolve stalls (profiles/r2_gemm_epilogue.md).
template <int BN, bool CONV, bool GEGLU, bool BRES, int EPI>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmD2, const __grid_constant__ CUtensorMap tmD1, const TcArgs args) {
  constexpr int MAX_STAGES = 8;
  constexpr int B_STAGE_BYTES = BN * BK * 2;
  constexpr int STAGE_BYTES = BRES ? A_STAGE_BYTES : A_STAGE_BYTES + B_STAGE_BYTES;
  constexpr int ACC_STRIDE = acc_stride(BN);
  constexpr int TMEM_COLS = 2 * ACC_STRIDE;
  constexpr uint32_t IDESC = make_idesc(BN);
  static_assert(BN % 16 == 0 && BN <= 256, "invalid UMMA N");
  static_assert(!(BRES && CONV), "resident weights are for small-K GEMMs");
  const int STAGES = BRES ? args.stages : num_stages(BN);

  extern __shared__ uint8_t smem_raw[];
  // 1 KB alignment by pointer arithmetic on the __shared__ array: an integer round trip makes every access through the
  // derived pointers a GENERIC load / store (the epilogue's bias-vector reads were LD.E.128 in the global-memory queue,
  // its top stall -- profiles/r2_gemm_epilogue.md)
  uint8_t* smem_base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* smem_b = smem_base;                                                        // BRES: num_k_blocks weight tiles
  uint8_t* smem = smem_base + (BRES ? args.num_k_blocks * B_STAGE_BYTES : 0);         // ring
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + MAX_STAGES;
  uint64_t* tmem_full = empty_bar + MAX_STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* b_full = tmem_empty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(b_full + 1);
  // Per-column epilogue vectors of the current tile (bias, LayerNorm column sums, row bias), one set per accumulator
  // stage: loaded by the epilogue warps BEFORE they wait for the accumulator, read back as shared-memory broadcasts.
  // (Read with __ldg inside the chunk loop they were the top stall of the small-K GEMMs: every 16-column chunk waited a
  // full L2 round trip for 64 bytes -- profiles/r2_gemm_epilogue.md.)
  float* s_vec = reinterpret_cast<float*>(full_bar) + 64;       // 256 bytes after the barriers; [2][3][256] floats
  uint8_t* s_out = reinterpret_cast<uint8_t*>(full_bar) + EPI_FIXED_BYTES;   // 1 KB aligned: TMA-store boxes, one per warp

  const int warp = uniform_warp_index(), lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(b_full, 1);
    mbar_init(&tmem_full[0], 1); mbar_init(&tmem_full[1], 1);
    mbar_init(&tmem_empty[0], 8); mbar_init(&tmem_empty[1], 8);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_prologue();   // everything above touched only shared / tensor memory; global reads and writes start below

  const int num_tiles = args.num_m_tiles * args.num_n_tiles;
  // Tile schedule.  Streaming: tile = blockIdx.x + i * gridDim.x, n fastest (the n-tiles of one m-block run on
  // neighbouring SMs at the same time, so A is fetched from HBM once).  BRES: the n-block is fixed per CTA
  // (gridDim.x is a multiple of num_n_tiles) and the CTA walks down the m-blocks.
  auto tile_at = [&](int i, int& m_blk, int& n_blk) -> bool {
    if (BRES) {
      const int m_stride = gridDim.x / args.num_n_tiles;
      n_blk = blockIdx.x % args.num_n_tiles;
      m_blk = blockIdx.x / args.num_n_tiles + i * m_stride;
      return m_blk < args.num_m_tiles;
    }
    const int tile = blockIdx.x + i * gridDim.x;
    m_blk = tile / args.num_n_tiles;
    n_blk = tile - m_blk * args.num_n_tiles;
    return tile < num_tiles;
  };

  if (warp == 0 && elect_one_sync()) {
    // ===================== TMA producer =====================
    int stage = 0;
    uint32_t phase = 0;
    int m_blk, n_blk;
    if (BRES && tile_at(0, m_blk, n_blk)) {
      mbar_expect_tx(b_full, args.num_k_blocks * B_STAGE_BYTES);
      for (int kb = 0; kb < args.num_k_blocks; ++kb) tma_load_2d(smem_b + kb * B_STAGE_BYTES, &tmB, b_full, kb * BK, n_blk * BN);
    }
    for (int it = 0; tile_at(it, m_blk, n_blk); ++it) {
      int cn = 0, ch = 0, cw = 0, par = 0;
      if (CONV) {
        int mb = m_blk;
        if (args.conv_mode == 2) { par = mb & 3; mb >>= 2; }
        if (args.pw) {
          const int tiles_w = args.W / args.pw, per_group = tiles_w * (args.H / args.ph);
          const int ng = mb / per_group, r = mb - ng * per_group;
          cn = ng * args.pf;
          ch = (r / tiles_w) * args.ph;
          cw = (r - (r / tiles_w) * tiles_w) * args.pw;
        } else {
          const int m0 = mb * BM, hw = args.H * args.W;
          cn = m0 / hw;
          const int rem = m0 - cn * hw;
          ch = rem / args.W;
          cw = rem - ch * args.W;
        }
      }
      // One pass of this loop per k-block is the producer's whole job: it has to stay well under the MMA time of a k-block
      // (320 clocks at BN = 160), so there is no division and no mode switch inside -- taps are an outer loop with the box
      // origin fixed per tap, the k coordinate of B just counts up (it was ~90 instructions with a software division before:
      // the conv at BN = 160 ran at the producer's pace, profiles/r2_mma_issue.md).
      int kb_b = par * args.num_k_blocks * BK;                      // k coordinate of the B (weight) box
      const int n0 = n_blk * BN;
      if (CONV) {
        const int taps = args.conv_mode == 2 ? 4 : 9, tw = args.conv_mode == 2 ? 2 : 3;
        const int sc = args.conv_mode == 1 ? 2 : 1;
        const int x0 = sc * cw - ((args.conv_mode == 2 && (par & 1)) ? 0 : 1);
        const int y0 = sc * ch - ((args.conv_mode == 2 && (par & 2)) ? 0 : 1);
        int tx = 0, ty = 0;
        for (int tap = 0; tap < taps; ++tap) {
          const int x = x0 + tx, y = y0 + ty;
          for (int cb = 0; cb < args.cin_blocks; ++cb) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* sa = smem + stage * STAGE_BYTES;
            mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
            tma_load_4d(sa, &tmA, &full_bar[stage], cb * BK, x, y, cn);
            tma_load_2d(sa + A_STAGE_BYTES, &tmB, &full_bar[stage], kb_b, n0);
            kb_b += BK;
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
          if (++tx == tw) { tx = 0; ++ty; }
        }
      } else {
        const int m0 = m_blk * BM;
        for (int kb = 0; kb < args.num_k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * STAGE_BYTES;
          mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
          tma_load_2d(sa, &tmA, &full_bar[stage], kb * BK, m0);
          if (!BRES) tma_load_2d(sa + A_STAGE_BYTES, &tmB, &full_bar[stage], kb * BK, n0);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
      if (!BRES) {
        for (int r = 0; r < args.res_kblocks; ++r) {        // [residual tile | identity] k-blocks (output rows m_blk * 128 ..)
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * STAGE_BYTES;
          mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
          tma_load_2d(sa, &tmD2, &full_bar[stage], n0 + r * BK, m_blk * BM);
          tma_load_2d(sa + A_STAGE_BYTES, &tmD1, &full_bar[stage], r * BK, 0);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1 && elect_one_sync()) {
    // ===================== MMA issuer =====================
    int stage = 0;
    uint32_t phase = 0;
    int m_blk, n_blk;
    if (BRES && tile_at(0, m_blk, n_blk)) mbar_wait(b_full, 0);
    for (int it = 0; tile_at(it, m_blk, n_blk); ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      tcgen05_fence_after();
      const uint32_t tmem_d = tmem_base + acc * ACC_STRIDE;
      const int kb_total = args.num_k_blocks + (BRES ? 0 : args.res_kblocks);
      for (int kb = 0; kb < kb_total; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tcgen05_fence_after();
        const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
        const uint32_t sb = BRES ? smem_u32(smem_b + kb * B_STAGE_BYTES) : sa + A_STAGE_BYTES;
        const uint64_t da = make_smem_desc(sa), db = make_smem_desc(sb);
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k) {
          // +32 bytes per UMMA_K step inside the 128B swizzle atom => +2 in the (addr >> 4) field
          umma_bf16(tmem_d, da + 2 * k, db + 2 * k, IDESC, (kb | k) != 0);
        }
        umma_commit(&empty_bar[stage]);   // frees the smem slot once these MMAs retire
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      umma_commit(&tmem_full[acc]);       // accumulator complete -> epilogue
    }
  } else if (warp >= 2) {
    // ===================== epilogue (8 warps) =====================
    // Warp w may only touch TMEM lanes 32*(w % 4)..; two warps share each lane group and split the tile's
    // 16-column chunks between them.  The residual of the whole row segment is prefetched before the
    // accumulator is waited for, so its global-load latency hides behind the main loop of this tile.
    const int lane_grp = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row_in_tile = lane_grp * 32 + lane;
    // GEGLU weights are row-interleaved in blocks of 16: tile columns [32p, 32p+16) hold "value" and
    // [32p+16, 32p+32) the matching "gate" columns of output columns [16p, 16p+16) of this tile.
    constexpr int OUT_COLS = GEGLU ? BN / 2 : BN;
    constexpr int NCH = OUT_COLS / 16;
    constexpr int CH0 = (NCH + 1) / 2;               // chunks of half 0; half 1 takes the rest
    const int c_begin = half ? CH0 : 0;
    const int c_count = half ? (NCH - CH0) : CH0;
    const bool wide = args.wide_io != 0;
    // Residual prefetch, one tile ahead: chunk slot i is consumed for tile `it` and immediately refilled with the
    // same chunk of tile `it + 1`, so every load has a whole tile period to land (for small K the epilogue, not the
    // main loop, is the critical path and nothing else would cover the HBM latency of these loads).
    Row32 resv[CH0];
    constexpr bool RES = (EPI == 2 || EPI == 4);        // lean epilogue adds a residual
    constexpr bool TMAST = (EPI == 4);                  // lean epilogue stores through TMA (two-chunk boxes + one-chunk tail)
    if (RES) {
#pragma unroll
      for (int i = 0; i < CH0; ++i)
#pragma unroll
        for (int e = 0; e < 8; ++e) resv[i].w[e] = 0u;
    }
    int m_blk, n_blk;
    // output row of this thread in M tile mb (>= M: none).  Patch tiles (CONV, args.pw != 0) are not consecutive rows.
    auto row_of = [&](int mb) -> int {
      int par = 0;
      if (CONV && args.conv_mode == 2) { par = mb & 3; mb >>= 2; }
      if (CONV && (args.pw || args.conv_mode == 2)) {
        int fr, hh, ww;
        if (args.pw) {
          const int tiles_w = args.W / args.pw, per_group = tiles_w * (args.H / args.ph);
          const int ng = mb / per_group, r = mb - ng * per_group;
          const int per_frame = args.pw * args.ph;
          const int dn = row_in_tile / per_frame, rr = row_in_tile - dn * per_frame;
          const int dh = rr / args.pw, dw = rr - dh * args.pw;
          fr = ng * args.pf + dn;
          hh = (r / tiles_w) * args.ph + dh;
          ww = (r - (r / tiles_w) * tiles_w) * args.pw + dw;
        } else {
          const int r = mb * BM + row_in_tile, hw = args.H * args.W;
          fr = r / hw;
          const int rem = r - fr * hw;
          hh = rem / args.W;
          ww = rem - hh * args.W;
        }
        if (fr >= args.n_frames) return args.M;
        if (args.conv_mode == 2) return (fr * 2 * args.H + 2 * hh + (par >> 1)) * 2 * args.W + 2 * ww + (par & 1);
        return (fr * args.H + hh) * args.W + ww;
      }
      return mb * BM + row_in_tile;
    };
    if (args.residual != nullptr && tile_at(0, m_blk, n_blk) && row_of(m_blk) < args.M) {
      const bf16* rp = args.residual + (int64_t)row_of(m_blk) * args.ldr + n_blk * OUT_COLS + c_begin * 16;
#pragma unroll
      for (int i = 0; i < CH0; ++i)
        if (i < c_count) resv[i] = ld_row32(rp + 16 * i, wide);
    }
    const int et = (int)threadIdx.x - 64;          // epilogue thread 0..255
    for (int it = 0; tile_at(it, m_blk, n_blk); ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int m = row_of(m_blk);
      const bool m_ok = m < args.M;
      float* sv = s_vec + acc * 768;
      const bool rb_staged = args.rowbias != nullptr && args.rb_tile_rows > 0;
      if (et < BN / 4) {
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        float4 b4 = args.bias ? __ldg(reinterpret_cast<const float4*>(args.bias + n_blk * BN) + et) : z;
        if (EPI != 0 && !GEGLU && rb_staged) {      // lean: bias + row bias of the tile's group in one vector
          int tile_row0 = m_blk * BM;
          if (CONV && args.conv_mode == 2) tile_row0 = (m_blk >> 2) * BM;
          int grp = tile_row0 / args.rb_tile_rows;
          if (args.rowbias_mod > 0) grp %= args.rowbias_mod;
          const float4 r4 = __ldg(reinterpret_cast<const float4*>(args.rowbias + (int64_t)grp * args.ld_rowbias + n_blk * BN) + et);
          b4.x += r4.x; b4.y += r4.y; b4.z += r4.z; b4.w += r4.w;
        }
        reinterpret_cast<float4*>(sv)[et] = b4;
        if (EPI == 0 && args.rowstats)
          reinterpret_cast<float4*>(sv + 256)[et] = __ldg(reinterpret_cast<const float4*>(args.colsum + n_blk * BN) + et);
        if (EPI == 0 && rb_staged && et < OUT_COLS / 4) {
          int tile_row0 = m_blk * BM;
          if (CONV && args.conv_mode == 2) tile_row0 = (m_blk >> 2) * BM;
          int grp = tile_row0 / args.rb_tile_rows;
          if (args.rowbias_mod > 0) grp %= args.rowbias_mod;
          reinterpret_cast<float4*>(sv + 512)[et] =
              __ldg(reinterpret_cast<const float4*>(args.rowbias + (int64_t)grp * args.ld_rowbias + n_blk * OUT_COLS) + et);
        }
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");     // the 8 epilogue warps: vectors of this tile are in place
      const int n_out0 = n_blk * OUT_COLS;
      const bool has_res = args.residual != nullptr && m_ok;
      const bf16* res_next = nullptr;
      {
        int mb2, nb2;
        if (args.residual != nullptr && tile_at(it + 1, mb2, nb2) && row_of(mb2) < args.M)
          res_next = args.residual + (int64_t)row_of(mb2) * args.ldr + nb2 * OUT_COLS + c_begin * 16;
      }
      bf16* drow = args.D + (int64_t)m * args.ldd;
      if (args.ex.direction != 0 && m_ok) drow = exchange_row_ptr<bf16>(args.ex, m);
      const float rs = (args.rowscale && m_ok ? args.rowscale[m] : 1.f) * args.alpha;
      const float* rb = nullptr;
      if (args.rowbias && m_ok && !rb_staged) {
        int grp = m / args.rows_per_group;
        if (args.rowbias_mod > 0) grp %= args.rowbias_mod;
        rb = args.rowbias + (int64_t)grp * args.ld_rowbias;
      }
      float ln_nmr = 0.f, ln_rstd = 1.f;      // (-rstd * mean, rstd) of this thread's row
      if (args.rowstats && m_ok) {
        const float2 st = __ldg(reinterpret_cast<const float2*>(args.rowstats) + m);
        ln_rstd = st.y; ln_nmr = -st.x * st.y;
      }
      mbar_wait(&tmem_full[acc], acc_phase);
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + acc * ACC_STRIDE + ((uint32_t)(lane_grp * 32) << 16);
      // TMEM loads run one chunk ahead (ping-pong registers): the tcgen05.ld of chunk i + 1 is in flight while chunk i goes
      // through bias / activation / residual / store.  (Issued and waited for inside the chunk, the load latency was the
      // top stall of the epilogue warps of every small-K GEMM: profiles/r2_gemm_epilogue.md.)
      if constexpr (EPI != 0) {
        // ---- lean epilogue: straight-line, the only predicates are the row guard on the store and the prefetch
        uint32_t rbuf[2][16], gbuf[GEGLU ? 2 : 1][16];
        {
          const int c0 = c_begin * 16;
          tmem_ld_x16(taddr + (GEGLU ? 2 * c0 : c0), rbuf[0]);
          if (GEGLU) tmem_ld_x16(taddr + 2 * c0 + 16, gbuf[0]);
        }
        bf16* dptr = drow + n_out0 + c_begin * 16;
        const uint32_t my_out = smem_u32(s_out) + (warp - 2) * 2048;
#pragma unroll
        for (int i = 0; i < CH0; ++i) {
          if (i < c_count) {
            const int c = (c_begin + i) * 16;
            uint32_t(&r)[16] = rbuf[i & 1];
            uint32_t(&g)[16] = gbuf[GEGLU ? (i & 1) : 0];
            float v[16];
            tmem_ld_wait();
            if (i + 1 < CH0 && i + 1 < c_count) {
              const int cn = c + 16;
              tmem_ld_x16(taddr + (GEGLU ? 2 * cn : cn), rbuf[(i + 1) & 1]);
              if (GEGLU) tmem_ld_x16(taddr + 2 * cn + 16, gbuf[GEGLU ? ((i + 1) & 1) : 0]);
            }
            pin16(r);
            if (GEGLU) pin16(g);
            float bv[16];
            lds16_f32(sv + (GEGLU ? 2 * c : c), bv);
            if (GEGLU) {
              float bg[16];
              lds16_f32(sv + 2 * c + 16, bg);
#pragma unroll
              for (int e = 0; e < 16; ++e)
                v[e] = (__uint_as_float(r[e]) + bv[e]) * gelu_logistic(__uint_as_float(g[e]) + bg[e]);
            } else {
#pragma unroll
              for (int e = 0; e < 16; ++e) v[e] = __uint_as_float(r[e]) + bv[e];
            }
            if (RES) {        // rows >= M add whatever the slot holds; their store is predicated off
              const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&resv[i].w[0]);
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                float2 f = __bfloat1622float2(h[e]);
                v[2 * e] += f.x; v[2 * e + 1] += f.y;
              }
            }
            if constexpr (TMAST) {
              // chunk -> swizzled box -> TMA store.  Two-chunk boxes (64-byte rows, 64B swizzle: 16-byte unit ^= (row >> 1)
              // & 3) while a partner chunk exists, else a one-chunk box (32-byte rows, 32B swizzle: unit ^= (row >> 2) & 1);
              // either way the 8 lanes of a store phase hit 8 different bank groups.
              const bool pair = (i & 1) || i + 1 < c_count;
              if (!pair || !(i & 1)) {          // opening a box: the previous store has finished reading the buffer
                if (lane == 0) tma_store_wait_read();
                __syncwarp();
              }
              uint32_t w[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
                w[e] = *reinterpret_cast<uint32_t*>(&h);
              }
              uint32_t a0, a1;
              if (pair) {
                const uint32_t sw = (lane >> 1) & 3, u = (i & 1) * 2;
                a0 = my_out + lane * 64 + ((u ^ sw) << 4);
                a1 = my_out + lane * 64 + (((u + 1) ^ sw) << 4);
              } else {
                const uint32_t sw = (lane >> 2) & 1;
                a0 = my_out + lane * 32 + (sw << 4);
                a1 = my_out + lane * 32 + ((1 ^ sw) << 4);
              }
              sts_v4(a0, w[0], w[1], w[2], w[3]);
              sts_v4(a1, w[4], w[5], w[6], w[7]);
              if (!pair || (i & 1)) {
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0)
                  tma_store_2d(pair ? &tmD2 : &tmD1, my_out, n_out0 + (pair ? c - 16 : c), m_blk * BM + lane_grp * 32);
              }
            } else {
              st_row32_if(m_ok, dptr + 16 * i, v);
            }
            if (RES) ld_row32_if(res_next != nullptr, resv[i], res_next + 16 * i);
          }
        }
      } else {
      uint32_t rbuf[2][16], gbuf[GEGLU ? 2 : 1][16];
      if (c_count > 0) {
        const int c0 = c_begin * 16;
        tmem_ld_x16(taddr + (GEGLU ? 2 * c0 : c0), rbuf[0]);
        if (GEGLU) tmem_ld_x16(taddr + 2 * c0 + 16, gbuf[0]);
      }
#pragma unroll
      for (int i = 0; i < CH0; ++i) {
        if (i < c_count) {
          const int c = (c_begin + i) * 16;
          uint32_t(&r)[16] = rbuf[i & 1];
          uint32_t(&g)[16] = gbuf[GEGLU ? (i & 1) : 0];
          float v[16];
          tmem_ld_wait();                       // chunk i has landed
          if (i + 1 < CH0 && i + 1 < c_count) {  // chunk i + 1 -> the other register set
            const int cn = (c_begin + i + 1) * 16;
            tmem_ld_x16(taddr + (GEGLU ? 2 * cn : cn), rbuf[(i + 1) & 1]);
            if (GEGLU) tmem_ld_x16(taddr + 2 * cn + 16, gbuf[GEGLU ? ((i + 1) & 1) : 0]);
          }
          if (GEGLU) {
            float bv[16], bg[16];
            lds16_f32(sv + 2 * c, bv);
            lds16_f32(sv + 2 * c + 16, bg);
            if (args.rowstats) {          // fused LayerNorm (warp-uniform branch): rstd * acc + (-rstd * mean) * colsum + bias
              float cv[16], cg[16];
              lds16_f32(sv + 256 + 2 * c, cv);
              lds16_f32(sv + 256 + 2 * c + 16, cg);
#pragma unroll
              for (int e = 0; e < 16; ++e) {
                r[e] = __float_as_uint(fmaf(ln_rstd, __uint_as_float(r[e]), fmaf(ln_nmr, cv[e], bv[e])));
                g[e] = __float_as_uint(fmaf(ln_rstd, __uint_as_float(g[e]), fmaf(ln_nmr, cg[e], bg[e])));
              }
            } else {
#pragma unroll
              for (int e = 0; e < 16; ++e) {
                r[e] = __float_as_uint(__uint_as_float(r[e]) + bv[e]);
                g[e] = __float_as_uint(__uint_as_float(g[e]) + bg[e]);
              }
            }
            if (args.act == 3) {          // A/B switch: one branch per chunk, not a select per element
#pragma unroll
              for (int e = 0; e < 16; ++e) v[e] = __uint_as_float(r[e]) * gelu_erf_fast(__uint_as_float(g[e]));
            } else {
#pragma unroll
              for (int e = 0; e < 16; ++e) v[e] = __uint_as_float(r[e]) * gelu_logistic(__uint_as_float(g[e]));
            }
          } else {
            float bv[16];
            lds16_f32(sv + c, bv);
            if (args.rowstats) {
              float cv[16];
              lds16_f32(sv + 256 + c, cv);
#pragma unroll
              for (int e = 0; e < 16; ++e) v[e] = fmaf(ln_rstd, __uint_as_float(r[e]), fmaf(ln_nmr, cv[e], bv[e]));
            } else {
#pragma unroll
              for (int e = 0; e < 16; ++e) v[e] = __uint_as_float(r[e]) + bv[e];
            }
          }
          if (m_ok) {
#pragma unroll
            for (int e = 0; e < 16; ++e) v[e] *= rs;
            if (rb_staged) {
              float rv[16];
              lds16_f32(sv + 512 + c, rv);
#pragma unroll
              for (int e = 0; e < 16; ++e) v[e] += rv[e];
            } else if (rb) {
              float rv[16];
              load16_f32(rb + n_out0 + c, rv);
#pragma unroll
              for (int e = 0; e < 16; ++e) v[e] += rv[e];
            }
            if (!GEGLU && args.act) {
#pragma unroll
              for (int e = 0; e < 16; ++e) v[e] = act_apply(v[e], args.act);
            }
            if (has_res) {
              const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&resv[i].w[0]);
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                float2 f = __bfloat1622float2(h[e]);
                v[2 * e] += f.x; v[2 * e + 1] += f.y;
              }
            }
            st_row32(drow + n_out0 + c, v, wide);
          }
          if (res_next) resv[i] = ld_row32(res_next + 16 * i, wide);
        }
      }
      }   // EPI == 0
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
    }
    if (EPI >= 3 && lane == 0) tma_store_wait_all();
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
  }
}

// ------------------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int resolve_encode(mmgt_ctx* ctx, EncodeTiledFn* fn) {
  if (!ctx->encode_tiled) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) {
      mmgt_set_error("cuTensorMapEncodeTiled not available (%s)", cudaGetErrorString(e));
      return MMGT_E_NODRIVER;
    }
    ctx->encode_tiled = p;
  }
  *fn = reinterpret_cast<EncodeTiledFn>(ctx->encode_tiled);
  return 0;
}

int make_map(mmgt_ctx* ctx, CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
             const uint32_t* box, const uint32_t* elem_strides = nullptr, CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
  EncodeTiledFn fn;
  int rc = resolve_encode(ctx, &fn);
  if (rc) return rc;
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  if (elem_strides)
    for (int i = 0; i < rank; ++i) estr[i] = elem_strides[i];
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_bytes, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    mmgt_set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d, dims %llu %llu, box %u %u)", (int)r, rank,
                   (unsigned long long)dims[0], (unsigned long long)dims[1], box[0], box[1]);
    return MMGT_E_INVALID;
  }
  return 0;
}

// Output maps of the TMA-store epilogue: (rows, cols) bf16 with row pitch ld; boxes of 32 rows x 32 / 16 columns.
struct OutMaps { CUtensorMap d2, d1; };
int make_out_maps(mmgt_ctx* ctx, OutMaps* o, const void* D, int64_t rows, int cols, int64_t ld) {
  uint64_t dims[2] = {(uint64_t)cols, (uint64_t)rows};
  uint64_t str[1] = {(uint64_t)ld * 2};
  uint32_t box2[2] = {32, 32}, box1[2] = {16, 32};
  int rc = make_map(ctx, &o->d2, D, 2, dims, str, box2, nullptr, CU_TENSOR_MAP_SWIZZLE_64B);
  if (rc) return rc;
  return make_map(ctx, &o->d1, D, 2, dims, str, box1, nullptr, CU_TENSOR_MAP_SWIZZLE_32B);
}

// Residual-through-MMA operands: d2 = the residual matrix as an A operand (boxes of 128 rows x 64 columns), d1 = the
// context's 256 x 256 identity as a B operand (boxes of bn rows x 64 columns).
int make_residual_maps(mmgt_ctx* ctx, OutMaps* o, const void* R, int64_t rows, int cols, int64_t ld, int bn) {
  {
    uint64_t dims[2] = {(uint64_t)cols, (uint64_t)rows};
    uint64_t str[1] = {(uint64_t)ld * 2};
    uint32_t box[2] = {BK, BM};
    int rc = make_map(ctx, &o->d2, R, 2, dims, str, box);
    if (rc) return rc;
  }
  uint64_t dims[2] = {256, 256};
  uint64_t str[1] = {256 * 2};
  uint32_t box[2] = {BK, (uint32_t)bn};
  return make_map(ctx, &o->d1, ctx->identity, 2, dims, str, box);
}

// The lean epilogue covers: bias, a row bias that is constant over each M tile (staged), GEGLU (logistic form), residual.
int pick_epi(const mmgt_ctx* ctx, const TcArgs& a, bool geglu) {
  if (!ctx->lean_epilogue || !a.wide_io || a.ex.direction != 0 || a.rowscale || a.alpha != 1.f || a.rowstats || a.act) return 0;
  if (a.rowbias && (geglu || a.rb_tile_rows <= 0)) return 0;
  if (geglu) return a.residual ? 0 : 1;
  return a.residual ? 2 : 1;
}

int pick_bn(int N) {
  const int cand[5] = {256, 160, 128, 64, 32};
  for (int i = 0; i < 5; ++i)
    if (N % cand[i] == 0) return cand[i];
  return 0;
}
// Largest tile width dividing N that still yields at least one tile per SM; small-M layers (16x16 / 8x8 latents)
// otherwise leave most of the 148 SMs idle.  Falls back to the width with the most tiles.
int pick_bn_for(int N, int M, int num_sms) {
  const int cand[5] = {256, 160, 128, 64, 32};
  const int m_tiles = (M + BM - 1) / BM;
  int best = 0;
  for (int i = 0; i < 5; ++i) {
    if (N % cand[i]) continue;
    best = cand[i];
    if (m_tiles * (N / cand[i]) >= num_sms) return cand[i];
    if (cand[i] <= 64) break;      // do not go below 64 just to add tiles (B-operand re-reads grow)
  }
  return best;
}

template <int BN, bool CONV, bool GEGLU, int EPI>
int launch_tc_epi(mmgt_ctx* ctx, const CUtensorMap& tmA, const CUtensorMap& tmB, const OutMaps& tmD, const TcArgs& a, cudaStream_t st) {
  constexpr int STAGES = num_stages(BN);
  constexpr int smem = STAGES * (A_STAGE_BYTES + BN * BK * 2) + 1024 + EPI_FIXED_BYTES + epi_out_bytes(2);
  static_assert(smem <= 232448, "streaming kernel exceeds the 227 KB opt-in");
  MMGT_CUDA_OK(mmgt_smem_optin(ctx, gemm_tc_kernel<BN, CONV, GEGLU, false, EPI>, smem));
  const int tiles = a.num_m_tiles * a.num_n_tiles;
  const int grid = tiles < ctx->num_sms ? tiles : ctx->num_sms;
  MMGT_CUDA_OK(mmgt_launch(ctx, gemm_tc_kernel<BN, CONV, GEGLU, false, EPI>, dim3(grid), dim3(NUM_THREADS), smem, st, tmA, tmB,
                           tmD.d2, tmD.d1, a));
  MMGT_LAUNCH_OK(ctx);
  return 0;
}
template <int BN, bool CONV, bool GEGLU>
int launch_tc(mmgt_ctx* ctx, const CUtensorMap& tmA, const CUtensorMap& tmB, const OutMaps& tmD, const TcArgs& a, cudaStream_t st) {
  if (a.epi == 1) return launch_tc_epi<BN, CONV, GEGLU, 1>(ctx, tmA, tmB, tmD, a, st);
  if (!GEGLU && a.epi == 2) return launch_tc_epi<BN, CONV, false, 2>(ctx, tmA, tmB, tmD, a, st);
  if (!GEGLU && a.epi == 4) return launch_tc_epi<BN, CONV, false, 4>(ctx, tmA, tmB, tmD, a, st);
  return launch_tc_epi<BN, CONV, GEGLU, 0>(ctx, tmA, tmB, tmD, a, st);
}

template <bool CONV>
int dispatch_tc(mmgt_ctx* ctx, int bn, bool geglu, const CUtensorMap& tmA, const CUtensorMap& tmB, const OutMaps& tmD,
                const TcArgs& a, cudaStream_t st) {
#define CASE(BN_)                                                              \
  case BN_:                                                                    \
    if (!CONV && geglu) return launch_tc<BN_, false, true>(ctx, tmA, tmB, tmD, a, st); \
    return launch_tc<BN_, CONV, false>(ctx, tmA, tmB, tmD, a, st);
  switch (bn) {
    CASE(256)
    CASE(160)
    CASE(128)
    CASE(64)
    CASE(32)
  }
#undef CASE
  mmgt_set_error("gemm_tc: no kernel for BN=%d", bn);
  return MMGT_E_UNSUPPORTED;
}

// ---- weight-stationary plan (BRES kernels)
constexpr int SMEM_OPTIN = 232448;          // 227 KB per CTA on sm_100
constexpr int BRES_OVERHEAD = 1024 + EPI_FIXED_BYTES;   // alignment slack + barriers + epilogue vectors

struct BresPlan { int bn, stages, grid, tma_store; };

// Resident weights pay off when the (BN x K) tile fits next to >= 3 A stages, at least 90 % of the SMs get a CTA
// (the grid must be a multiple of the n-tile count) and every CTA has a few m-tiles to amortise the weight load.
bool plan_bres(int M, int N, int K, bool geglu, int num_sms, bool tma_store_ok, BresPlan* out) {
  const int kblocks = (K + BK - 1) / BK;
  const int m_tiles = (M + BM - 1) / BM;
  // BN = 128 (the only width whose K = 640 weight tile fits) is not offered: at 128 x 128 the MMA reads as many operand
  // bytes from shared memory per flop as the ring can deliver, three A stages are all that is left, and the streaming
  // kernel at BN = 160 measured 19-33 % faster on every K = 640 shape (profiles/r2_bres_sweep.txt).
  const int cand[3] = {256, 240, 160};
  for (int i = 0; i < 3; ++i) {
    const int bn = cand[i];
    if (N % bn || (geglu && bn % 32)) continue;
    const int n_tiles = N / bn;
    const int b_bytes = kblocks * bn * BK * 2;
    // staging for the TMA-store epilogue: two-chunk boxes when >= 3 A stages still fit, else one-chunk boxes
    int tma_store = (tma_store_ok && !geglu) ? 2 : 0;
    if (tma_store == 2 && (SMEM_OPTIN - BRES_OVERHEAD - epi_out_bytes(2) - b_bytes) / A_STAGE_BYTES < 3) tma_store = 0;
    int stages = (SMEM_OPTIN - BRES_OVERHEAD - epi_out_bytes(tma_store) - b_bytes) / A_STAGE_BYTES;
    if (stages > 8) stages = 8;
    if (stages < 3 || n_tiles > num_sms) continue;
    const int grid = (num_sms / n_tiles) * n_tiles;
    if (grid * 10 < num_sms * 9) continue;
    if (m_tiles < 4 * (grid / n_tiles)) continue;
    out->bn = bn; out->stages = stages; out->grid = grid; out->tma_store = tma_store;
    return true;
  }
  return false;
}

template <int BN, bool GEGLU, int EPI>
int launch_bres_epi(mmgt_ctx* ctx, const CUtensorMap& tmA, const CUtensorMap& tmB, const OutMaps& tmD, const TcArgs& a, int grid,
                    cudaStream_t st) {
  MMGT_CUDA_OK(mmgt_smem_optin(ctx, gemm_tc_kernel<BN, false, GEGLU, true, EPI>, SMEM_OPTIN));
  const int smem = BRES_OVERHEAD + epi_out_bytes(a.tma_store) + a.num_k_blocks * BN * BK * 2 + a.stages * A_STAGE_BYTES;
  MMGT_CUDA_OK(mmgt_launch(ctx, gemm_tc_kernel<BN, false, GEGLU, true, EPI>, dim3(grid), dim3(NUM_THREADS), smem, st, tmA, tmB,
                           tmD.d2, tmD.d1, a));
  MMGT_LAUNCH_OK(ctx);
  return 0;
}
template <int BN, bool GEGLU>
int launch_bres(mmgt_ctx* ctx, const CUtensorMap& tmA, const CUtensorMap& tmB, const OutMaps& tmD, const TcArgs& a, int grid,
                cudaStream_t st) {
  if (a.epi == 1) return launch_bres_epi<BN, GEGLU, 1>(ctx, tmA, tmB, tmD, a, grid, st);
  if (!GEGLU && a.epi == 2) return launch_bres_epi<BN, false, 2>(ctx, tmA, tmB, tmD, a, grid, st);
  if (!GEGLU && a.epi == 4) return launch_bres_epi<BN, false, 4>(ctx, tmA, tmB, tmD, a, grid, st);
  return launch_bres_epi<BN, GEGLU, 0>(ctx, tmA, tmB, tmD, a, grid, st);
}

int dispatch_bres(mmgt_ctx* ctx, const BresPlan& pl, bool geglu, const CUtensorMap& tmA, const CUtensorMap& tmB,
                  const OutMaps& tmD, const TcArgs& a, cudaStream_t st) {
#define CASE(BN_)                                                                    \
  case BN_:                                                                          \
    if (geglu) return launch_bres<BN_, true>(ctx, tmA, tmB, tmD, a, pl.grid, st);    \
    return launch_bres<BN_, false>(ctx, tmA, tmB, tmD, a, pl.grid, st);
  switch (pl.bn) {
    CASE(256)
    CASE(160)
    case 240: return launch_bres<240, false>(ctx, tmA, tmB, tmD, a, pl.grid, st);
  }
#undef CASE
  mmgt_set_error("gemm_tc: no resident-weight kernel for BN=%d", pl.bn);
  return MMGT_E_UNSUPPORTED;
}

}  // namespace

extern "C" int mmgt_gemm_tc_block_n(int N) { return pick_bn(N); }

bool mmgt_gemm_tc_supported(const mmgt_ctx* ctx, const mmgt_gemm_params* p) {
  (void)ctx;
  if (p->dtype != MMGT_BF16 || p->out_f32) return false;
  const int bn = pick_bn(p->N);
  if (!bn) return false;
  if (p->geglu_block && p->geglu_block != 16) return false;
  if (p->K % 8 || p->lda % 8 || p->ldw % 8) return false;
  if (!aligned16(p->A) || !aligned16(p->W)) return false;
  const int n_out = p->geglu_block ? p->N / 2 : p->N;
  if (p->exchange) {
    if (p->exchange->ld % 8 || p->exchange->ld < n_out) return false;
  } else if (!aligned16(p->D) || p->ldd % 8) {
    return false;
  }
  if (n_out % 16) return false;
  if (p->residual && (!aligned16(p->residual) || p->ldr % 8)) return false;
  if ((p->bias && !aligned16(p->bias)) || (p->rowbias && !aligned16(p->rowbias))) return false;   // float4 epilogue loads
  if (p->rowbias && p->ld_rowbias % 4) return false;
  if (p->rowstats && (!aligned16(p->colsum) || (reinterpret_cast<uintptr_t>(p->rowstats) & 7u))) return false;
  if (p->geglu_block && p->act) return false;
  if (p->M < 1) return false;
  return true;
}

int mmgt_gemm_tc(mmgt_ctx* ctx, const mmgt_gemm_params* p, cudaStream_t st) {
  const bool geglu = p->geglu_block != 0;
  // ---- epilogue options first: they decide the kernel family
  TcArgs a{};
  a.M = p->M;
  a.N_out = geglu ? p->N / 2 : p->N;
  a.num_m_tiles = (p->M + BM - 1) / BM;
  a.num_k_blocks = (p->K + BK - 1) / BK;
  a.bias = p->bias; a.rowscale = p->rowscale; a.rowbias = p->rowbias;
  a.residual = (const bf16*)p->residual; a.D = (bf16*)p->D;
  a.ldd = p->ldd; a.ldr = p->ldr; a.rows_per_group = p->rows_per_group > 0 ? p->rows_per_group : 1; a.alpha = p->alpha;
  a.ld_rowbias = p->ld_rowbias ? p->ld_rowbias : a.N_out; a.rowbias_mod = p->rowbias_mod;
  a.rowstats = p->rowstats; a.colsum = p->colsum;
  a.act = geglu ? (ctx->geglu_exact ? 3 : 0) : p->act;
  a.rb_tile_rows = (p->rowbias && a.rows_per_group % BM == 0) ? a.rows_per_group : 0;
  a.wide_io = aligned32(p->D) && p->ldd % 16 == 0 && (!p->residual || (aligned32(p->residual) && p->ldr % 16 == 0));
  if (p->exchange) {
    int rc = mmgt_row_exchange_check(p->exchange, p->M, "gemm");
    if (rc) return rc;
    a.ex = *p->exchange;
    bool w = p->exchange->ld % 16 == 0 && (!p->residual || (aligned32(p->residual) && p->ldr % 16 == 0));
    for (int s = 0; s < p->exchange->k; ++s) w = w && aligned32(p->exchange->peer_base[s]);
    a.wide_io = w;
  }
  a.epi = pick_epi(ctx, a, geglu);
  // residual epilogue: through the tensor cores on the streaming kernel where allowed, else TMA stores + per-lane loads
  // Measured (profiles/r2_bres_sweep_residual.txt): the extra k-blocks pay only where the epilogue, not the main loop, paces
  // the kernel -- 320x320 + residual 54.3 -> 47.2 us, but 640x640 97.2 -> 101.4, 640x2560 78.8 -> 82.9, 1280x1280 27.6 -> 30.7 us.
  const bool res_mma = ctx->residual_mma && a.epi == 2 && ctx->identity != nullptr && p->K <= 512;
  BresPlan pl{};
  const bool tma_store_ok = ctx->tma_store && a.epi == 2 && !res_mma;      // see the note at TcArgs::tma_store
  const bool bres = ctx->use_bres && !res_mma && plan_bres(p->M, p->N, p->K, geglu, ctx->num_sms, tma_store_ok, &pl);
  const int bn = bres ? pl.bn : pick_bn_for(p->N, p->M, ctx->num_sms);
  a.num_n_tiles = p->N / bn;
  CUtensorMap tmA, tmB;
  {
    uint64_t dims[2] = {(uint64_t)p->K, (uint64_t)p->M};
    uint64_t str[1] = {(uint64_t)p->lda * 2};
    uint32_t box[2] = {BK, BM};
    int rc = make_map(ctx, &tmA, p->A, 2, dims, str, box);
    if (rc) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)p->K, (uint64_t)p->N};
    uint64_t str[1] = {(uint64_t)p->ldw * 2};
    uint32_t box[2] = {BK, (uint32_t)bn};
    int rc = make_map(ctx, &tmB, p->W, 2, dims, str, box);
    if (rc) return rc;
  }
  OutMaps tmD;
  tmD.d2 = tmA; tmD.d1 = tmA;      // unused unless set below
  // Measured (profiles/r2_ab_flags.md): the TMA-store path wins where the epilogue also LOADS a residual row per lane
  // (320x320 + residual 66.6 -> 54.3 us: half the LSU work), and loses slightly where it only stores (960x320 70.6 ->
  // 74.7 us: the staging traffic competes with the MMA operand reads for shared-memory bandwidth); one-chunk boxes alone
  // serialise on the single buffer.  So: residual epilogues with room for two-chunk boxes only.
  a.tma_store = !tma_store_ok ? 0 : bres ? pl.tma_store : 2;
  if (a.tma_store != 2) a.tma_store = 0;
  if (res_mma) {
    int rc = make_residual_maps(ctx, &tmD, p->residual, p->M, a.N_out, p->ldr, bn);
    if (rc) return rc;
    a.res_kblocks = (bn + BK - 1) / BK;
    a.residual = nullptr;
    a.epi = 1;
  } else if (a.tma_store) {
    int rc = make_out_maps(ctx, &tmD, p->D, p->M, a.N_out, p->ldd);
    if (rc) return rc;
    a.epi = 4;
  }
  if (bres) {
    a.stages = pl.stages;      // planned with the staging boxes reserved; a launch without them just leaves them unused
    return dispatch_bres(ctx, pl, geglu, tmA, tmB, tmD, a, st);
  }
  return dispatch_tc<false>(ctx, bn, geglu, tmA, tmB, tmD, a, st);
}

static bool conv_box(int W, int H, uint32_t* bw, uint32_t* bh, uint32_t* bn_frames) {
  if (W >= BM) {
    if (W % BM) return false;
    *bw = BM; *bh = 1; *bn_frames = 1;
    return true;
  }
  if (BM % W) return false;
  int rows = BM / W;
  if (rows <= H) {
    if (H % rows) return false;
    *bw = W; *bh = rows; *bn_frames = 1;
    return true;
  }
  if (rows % H) return false;
  *bw = W; *bh = H; *bn_frames = rows / H;
  return true;
}

// Patch tile for widths that conv_box cannot cover with consecutive rows: the widest pw | W with 128 % pw == 0, then
// the tallest ph | H with ph | 128 / pw; the rest of the 128 rows are consecutive frames.
static bool conv_patch(int W, int H, uint32_t* pw, uint32_t* ph, uint32_t* pf) {
  for (int w = 64; w >= 1; w >>= 1) {
    if (W % w) continue;
    const int rest = BM / w;
    for (int h = rest; h >= 1; h >>= 1) {
      if (H % h) continue;
      if (rest / h > 256) return false;
      *pw = w; *ph = h; *pf = rest / h;
      return true;
    }
  }
  return false;
}

// conv_mode: 0 = 3x3 stride 1, 1 = 3x3 stride 2 (TMA traversal stride), 2 = nearest x2 upsample + 3x3 as four 2x2-tap
// sub-pixel convolutions (needs p->w_subpixel).  Tile space: output image for 0 / 1, input image for 2.
static int conv_mode_of(const mmgt_conv3x3_params* p) { return p->upsample2x ? 2 : (p->stride == 2 ? 1 : 0); }
static void conv_tile_space(const mmgt_conv3x3_params* p, int* Ht, int* Wt) {
  if (p->stride == 2 && !p->upsample2x) { *Ht = p->H / 2; *Wt = p->W / 2; }
  else { *Ht = p->H; *Wt = p->W; }
}

bool mmgt_conv3x3_tc_supported(const mmgt_ctx* ctx, const mmgt_conv3x3_params* p) {
  (void)ctx;
  if (p->dtype != MMGT_BF16) return false;
  if (p->upsample2x && (p->stride != 1 || !p->w_subpixel || !aligned16(p->w_subpixel))) return false;
  if (p->stride == 2 && ((p->H | p->W) & 1)) return false;
  if (p->Cin % BK || !pick_bn(p->Cout) || p->Cout % 16) return false;
  int Ht, Wt;
  conv_tile_space(p, &Ht, &Wt);
  uint32_t bw, bh, bf;
  if (!conv_box(Wt, Ht, &bw, &bh, &bf) && !conv_patch(Wt, Ht, &bw, &bh, &bf)) return false;
  if (p->stride == 2 && (2 * bw > 256 || 2 * bh > 256)) return false;
  if (!aligned16(p->x) || !aligned16(p->w) || !aligned16(p->y) || (p->residual && !aligned16(p->residual))) return false;
  if ((p->bias && !aligned16(p->bias)) || (p->rowbias && !aligned16(p->rowbias))) return false;
  if (p->rowbias && p->ld_rowbias % 4) return false;
  return true;
}

int mmgt_conv3x3_tc(mmgt_ctx* ctx, const mmgt_conv3x3_params* p, cudaStream_t st) {
  const int mode = conv_mode_of(p);
  int Ht, Wt;
  conv_tile_space(p, &Ht, &Wt);
  const int Ho = mode == 2 ? 2 * p->H : Ht, Wo = mode == 2 ? 2 * p->W : Wt;
  const int bn = pick_bn_for(p->Cout, p->N * Ho * Wo, ctx->num_sms);
  uint32_t bw, bh, bf;
  const bool patch = !conv_box(Wt, Ht, &bw, &bh, &bf);
  if (patch) conv_patch(Wt, Ht, &bw, &bh, &bf);
  const int taps = mode == 2 ? 4 : 9;
  CUtensorMap tmA, tmB;
  {
    uint64_t dims[4] = {(uint64_t)p->Cin, (uint64_t)p->W, (uint64_t)p->H, (uint64_t)p->N};
    uint64_t str[3] = {(uint64_t)p->Cin * 2, (uint64_t)p->W * p->Cin * 2, (uint64_t)p->H * p->W * p->Cin * 2};
    uint32_t box[4] = {BK, bw, bh, bf};
    uint32_t estr[4] = {1, 1, 1, 1};
    if (mode == 1) { box[1] = 2 * bw; box[2] = 2 * bh; estr[1] = 2; estr[2] = 2; }
    int rc = make_map(ctx, &tmA, p->x, 4, dims, str, box, estr);
    if (rc) return rc;
  }
  {
    const uint64_t kw = (uint64_t)(mode == 2 ? 16 : 9) * p->Cin;      // mode 2: (Cout, 4 parities, 2, 2, Cin)
    uint64_t dims[2] = {kw, (uint64_t)p->Cout};
    uint64_t str[1] = {kw * 2};
    uint32_t box[2] = {BK, (uint32_t)bn};
    int rc = make_map(ctx, &tmB, mode == 2 ? p->w_subpixel : p->w, 2, dims, str, box);
    if (rc) return rc;
  }
  TcArgs a{};
  a.M = p->N * Ho * Wo;
  a.N_out = p->Cout;
  const int m_tile_space = p->N * Ht * Wt;
  a.num_m_tiles = (m_tile_space + BM - 1) / BM;
  a.n_frames = p->N;
  if (patch) {
    a.pw = (int)bw; a.ph = (int)bh; a.pf = (int)bf;
    a.num_m_tiles = ((p->N + (int)bf - 1) / (int)bf) * (Ht / (int)bh) * (Wt / (int)bw);
  }
  if (mode == 2) a.num_m_tiles *= 4;
  a.num_n_tiles = p->Cout / bn;
  a.cin_blocks = p->Cin / BK;
  a.num_k_blocks = taps * a.cin_blocks;
  a.conv_mode = mode;
  a.bias = p->bias; a.rowscale = nullptr; a.rowbias = p->rowbias;
  a.residual = (const bf16*)p->residual; a.D = (bf16*)p->y;
  a.ldd = p->Cout; a.ldr = p->Cout;
  a.wide_io = aligned32(p->y) && p->Cout % 16 == 0 && (!p->residual || aligned32(p->residual));
  a.rows_per_group = p->rowbias ? p->frames_per_group * Ho * Wo : 1;
  a.ld_rowbias = p->ld_rowbias ? p->ld_rowbias : p->Cout;
  a.act = p->act;
  a.alpha = 1.f;
  {   // row-bias group in TILE-SPACE rows (mode 2 tiles walk the input image)
    const int tile_rows_per_group = p->frames_per_group * Ht * Wt;
    a.rb_tile_rows = (p->rowbias && !patch && tile_rows_per_group % BM == 0) ? tile_rows_per_group : 0;
  }
  a.H = Ht; a.W = Wt;
  a.epi = pick_epi(ctx, a, false);
  OutMaps tmD;
  tmD.d2 = tmA; tmD.d1 = tmA;
  const bool row_tiles = !patch && mode != 2;                // tiles of 128 consecutive output rows
  if (a.epi == 2 && row_tiles && ctx->residual_mma && ctx->identity && taps * p->Cin <= 512) {     // never: K >= 576 (see mmgt_gemm_tc)
    int rc = make_residual_maps(ctx, &tmD, p->residual, a.M, p->Cout, p->Cout, bn);
    if (rc) return rc;
    a.res_kblocks = (bn + BK - 1) / BK;
    a.residual = nullptr;
    a.epi = 1;
  } else if (a.epi == 2 && row_tiles && ctx->tma_store) {
    int rc = make_out_maps(ctx, &tmD, p->y, a.M, p->Cout, p->Cout);
    if (rc) return rc;
    a.tma_store = 2;
    a.epi = 4;
  }
  return dispatch_tc<true>(ctx, bn, false, tmA, tmB, tmD, a, st);
}
