// Flash-style attention on tcgen05 / TMEM, fed by TMA (bf16 operands, fp32 softmax and accumulation).
//
//   out[n, q, h, :] = softmax_k( q . k * scale ) v     over keys [ per-frame segment ; optional shared segment ]
//
// Kernels in this file (mmgt_attention_tc picks one):
//   attention_tc_q256_kernel<LMMA, POLY>     head dim <= 64, Lq > 128 (default): one CTA = (frame, head, 256 queries), two
//                                            softmax groups on different query tiles, one MMA-issuing warp each, S / O / P in TMEM
//   attention_tc_kernel<DCH, BN, PACKED>     head dim 80 / 160 (DCH = 2 / 3) and short query sets: one CTA = (frame, head,
//                                            128 queries), the two softmax groups take alternate key tiles, merge at the end
//   attention_tc64_kernel, attention_tc_persist_kernel   measured-slower A/B forms of the 128-query kernel (flags 9, 14)
// Common structure of a CTA:
//   warp 0 (one lane)  TMA producer: Q once, then K / V tiles through two smem rings
//   warp 1 (one lane)  MMA issuer:   S = Q K_j^T   (128 x BN x d_pad, accumulator in TMEM, double buffered)
//                                    O += P_j V_j  (128 x d_pad x BN, V MN-major; P from TMEM at head dim <= 64, else smem)
//   warps 2..9         softmax: one thread per query row (= TMEM lane): tcgen05.ld S -> online max / exp2 -> bf16 P;
//                      lazy rescale of O in TMEM (only when the running max moved by > 2^8); final 1/l scaling and the
//                      global store of O.
// Head dims that are not multiples of 64 (40, 80, 160) need no padded tensors: Q/K/V are described to TMA as
// (d, heads, rows) and a 64-wide box on the d axis is zero-filled past d, which pads every head to 64/128/192.
// The second key segment implements ReferenceNet feature injection: frames whose seg2 index is -1 (the CFG
// "uncond" half, mutual_self_attention.py:168-188) simply stop after the first segment.
#include <cuda.h>

#include <algorithm>

#include "common.cuh"

namespace {

constexpr int BQ = 128;
constexpr int NUM_THREADS = 320;   // TMA warp, MMA warp, 2 x 4 softmax warps
// K / V ring depths per head-dim class (DCH = ceil(d / 64)): as deep as shared memory allows.  K is requested when
// QK(j) retires and V when PV(j) retires; with two stages the request precedes the use by ~1.5 tile periods, less
// than a TMA round trip, and the softmax groups end up waiting for S.
__host__ __device__ constexpr int k_stages(int dch) { return dch == 1 ? 4 : dch == 2 ? 2 : 3; }
__host__ __device__ constexpr int v_stages(int dch) { return dch == 1 ? 3 : 2; }
constexpr float RESCALE_THRESHOLD = 8.0f;   // log2 units

struct AttnArgs {
  int N, Lq, Lk, Lk2, heads, d;
  const int32_t* seg2_index;
  int has_seg2;
  bf16* out;
  int64_t ldo;
  float scale_log2e;
};

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
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// K-major operand, 128B swizzle, 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t addr) {
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}
// MN-major operand (V: keys x d, d contiguous), 128B swizzle: 64-element d chunks `lbo_bytes` apart,
// 8-key groups 1024 B apart
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t addr, uint32_t lbo_bytes) {
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__host__ __device__ constexpr uint32_t idesc_bf16(int n, bool b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (b_mn_major ? (1u << 16) : 0u) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(BQ >> 4) << 24);
}
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
// A operand from tensor memory (M = 128 lanes x K 16-bit values, two per 32-bit column), B from shared memory
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// Orders the consumers of r[] after the tcgen05.wait::ld that precedes it (the loads are asynchronous and their
// destination registers must not be read or moved before the wait).
__device__ __forceinline__ void pin32(uint32_t (&r)[32]) {
  asm volatile("" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
               "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]));
  asm volatile("" : "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
               "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31]));
}
__device__ __forceinline__ float ex2_approx(float x) {   // single MUFU.EX2, flush-to-zero (exp2f adds denormal fix-ups)
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
// Packed fp32 pairs (FFMA2 / FADD2 on sm_100): the softmax inner loop issues one instruction per two scores for the
// scale-and-subtract and for the row sum.  Same IEEE operations as the scalar forms, so results do not change.
__device__ __forceinline__ uint64_t pack_f32x2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack_f32x2(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t fma_f32x2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t add_f32x2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
// p[0..3] = 2^(s[i..i+3] * c - mc), packed as two bf16 pairs into pk[0..1], added lane-wise to the four partial sums
template <bool PACKED>
__device__ __forceinline__ void exp4_pack_sum(const uint32_t* sc, float c, float mc, uint32_t* pk, float (&lsum)[4]) {
  float p0, p1, p2, p3;
  if constexpr (PACKED) {
    const uint64_t c2 = pack_f32x2(c, c), nm2 = pack_f32x2(-mc, -mc);
    float x0, x1, x2, x3;
    unpack_f32x2(fma_f32x2(pack_f32x2(__uint_as_float(sc[0]), __uint_as_float(sc[1])), c2, nm2), x0, x1);
    unpack_f32x2(fma_f32x2(pack_f32x2(__uint_as_float(sc[2]), __uint_as_float(sc[3])), c2, nm2), x2, x3);
    p0 = ex2_approx(x0); p1 = ex2_approx(x1); p2 = ex2_approx(x2); p3 = ex2_approx(x3);
    unpack_f32x2(add_f32x2(pack_f32x2(lsum[0], lsum[1]), pack_f32x2(p0, p1)), lsum[0], lsum[1]);
    unpack_f32x2(add_f32x2(pack_f32x2(lsum[2], lsum[3]), pack_f32x2(p2, p3)), lsum[2], lsum[3]);
  } else {
    p0 = ex2_approx(fmaf(__uint_as_float(sc[0]), c, -mc));
    p1 = ex2_approx(fmaf(__uint_as_float(sc[1]), c, -mc));
    p2 = ex2_approx(fmaf(__uint_as_float(sc[2]), c, -mc));
    p3 = ex2_approx(fmaf(__uint_as_float(sc[3]), c, -mc));
    lsum[0] += p0; lsum[1] += p1; lsum[2] += p2; lsum[3] += p3;
  }
  pk[0] = pack_bf16(p0, p1);
  pk[1] = pack_bf16(p2, p3);
}

// 2^x for a pair of scores on the FMA pipe instead of MUFU (x <= 8; anything below -126 comes out as 2^-126): round x to
// the nearest integer n with the 1.5 * 2^23 trick, 2^(x - n) by a degree-3 minimax polynomial on [-0.5, 0.5] (relative error
// 7.5e-5, 50 times below the bf16 rounding P goes through), n added into the exponent field.  Seven packed instructions,
// two FMNMX and two LEA per pair against two MUFU.EX2: MUFU is the pipe that bounds head-dim-40 attention (16 results per
// clock and SM), the FMA pipe has slots to spare.
__device__ __forceinline__ void exp2_pair_fma(uint64_t x01, float& p0, float& p1) {
  float x0, x1;
  unpack_f32x2(x01, x0, x1);
  const uint64_t xc = pack_f32x2(fmaxf(x0, -126.f), fmaxf(x1, -126.f));
  const uint64_t xr = add_f32x2(xc, pack_f32x2(12582912.f, 12582912.f));
  const uint64_t nn = add_f32x2(xr, pack_f32x2(-12582912.f, -12582912.f));
  const uint64_t f = fma_f32x2(nn, pack_f32x2(-1.f, -1.f), xc);
  uint64_t p = fma_f32x2(pack_f32x2(0.05517149344086647f, 0.05517149344086647f), f, pack_f32x2(0.24261093139648438f, 0.24261093139648438f));
  p = fma_f32x2(p, f, pack_f32x2(0.6932609677314758f, 0.6932609677314758f));
  p = fma_f32x2(p, f, pack_f32x2(0.9999280571937561f, 0.9999280571937561f));
  float q0, q1, r0, r1;
  unpack_f32x2(p, q0, q1);
  unpack_f32x2(xr, r0, r1);
  p0 = __int_as_float(__float_as_int(q0) + (__float_as_int(r0) << 23));
  p1 = __int_as_float(__float_as_int(q1) + (__float_as_int(r1) << 23));
}
// Eight scores: p = 2^(s * c - mc) as four bf16 pairs in pk[0..3], added lane-wise to the partial sums.  POLY of the four
// pairs (0, 1 or 2) take the FMA-pipe exponential, the rest MUFU.EX2.
// SUM = false: the row sums come from the tensor cores (P x ones), nothing is added here.
template <int POLY, bool SUM>
__device__ __forceinline__ void exp8_pack_sum(const uint32_t* sc, float c, float mc, uint32_t* pk, float (&lsum)[4]) {
  const uint64_t c2 = pack_f32x2(c, c), nm2 = pack_f32x2(-mc, -mc);
  uint64_t l01 = 0, l23 = 0;
  if constexpr (SUM) { l01 = pack_f32x2(lsum[0], lsum[1]); l23 = pack_f32x2(lsum[2], lsum[3]); }
#pragma unroll
  for (int pr = 0; pr < 4; ++pr) {
    const uint64_t x = fma_f32x2(pack_f32x2(__uint_as_float(sc[2 * pr]), __uint_as_float(sc[2 * pr + 1])), c2, nm2);
    float p0, p1;
    if ((POLY >= 1 && pr == 3) || (POLY >= 2 && pr == 1)) {
      exp2_pair_fma(x, p0, p1);
    } else {
      float x0, x1;
      unpack_f32x2(x, x0, x1);
      p0 = ex2_approx(x0);
      p1 = ex2_approx(x1);
    }
    pk[pr] = pack_bf16(p0, p1);
    if constexpr (SUM) {
      if (pr & 1) l23 = add_f32x2(l23, pack_f32x2(p0, p1));
      else l01 = add_f32x2(l01, pack_f32x2(p0, p1));
    }
  }
  if constexpr (SUM) {
    unpack_f32x2(l01, lsum[0], lsum[1]);
    unpack_f32x2(l23, lsum[2], lsum[3]);
  }
}

// DCH = ceil(d / 64) head-dim chunks (d_pad = 64 * DCH); BN = keys per tile.
// Two independent softmax groups (warps 2-5 and 6-9) take alternate key tiles, each with its own running
// maximum / row sum and its own O accumulator in TMEM (split-KV inside the CTA); the halves are merged in the
// epilogue.  With two warps per scheduler the TMEM-load, MUFU and shared-store phases of the groups overlap.
template <int DCH, int BN, bool PACKED>
__global__ void __launch_bounds__(NUM_THREADS, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmK2,
                    const __grid_constant__ CUtensorMap tmV2, const AttnArgs args) {
  constexpr int DPAD = 64 * DCH;
  constexpr int KS = k_stages(DCH), VS = v_stages(DCH);
  constexpr int Q_BYTES = BQ * DPAD * 2;
  constexpr int KV_BYTES = BN * DPAD * 2;        // one K (or V) stage
  constexpr int KV_CHUNK = BN * 128;             // bytes of one 64-wide d chunk of a K/V stage
  // P (the bf16 probabilities, A operand of P V) lives in TMEM when the 512 columns allow it (d <= 64): the softmax
  // warps write it with tcgen05.st and the MMA reads it in place, instead of a round trip through 128B-swizzled
  // shared memory.  Per 128 x 128 tile that removes 64 KB of shared-memory traffic (out of ~130 KB, at 128 B/clk a
  // ~1000 clk resource just like MUFU and the TMEM read port) and the generic->async proxy fence per tile.
  constexpr bool P_TMEM = (2 * BN + 2 * DPAD + BN) <= 512;
  constexpr int P_BYTES = P_TMEM ? 0 : BQ * BN * 2;   // one P buffer in shared memory (per softmax group)
  constexpr int TM_S = 0, TM_O = 2 * BN;         // S buffers at 0 / BN, O accumulators at 2BN / 2BN + DPAD
  constexpr int TM_P = TM_O + 2 * DPAD;          // P_TMEM: packed bf16 pairs, BN / 2 columns per group
  constexpr uint32_t IDESC_QK = idesc_bf16(BN, false);
  // Only ceil(d / 16) k-steps of Q K^T and round_up(d, 16) columns of P V are real: the rest of the 64-wide
  // TMA box is zero fill (d = 40 -> 3 of 4 k-steps, N = 48 of 64).
  const int qk_steps = (args.d + 15) >> 4;
  const uint32_t IDESC_PV = idesc_bf16(qk_steps << 4, true);
  static_assert(TM_O + 2 * DPAD + (P_TMEM ? BN : 0) <= 512, "TMEM budget");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + Q_BYTES;
  uint8_t* sV = sK + KS * KV_BYTES;
  uint8_t* sP = sV + VS * KV_BYTES;
  float* stats = reinterpret_cast<float*>(sP + 2 * P_BYTES);          // [128][2]: (m_ref, l) of group 1
  uint64_t* bars = reinterpret_cast<uint64_t*>(stats + 2 * BQ);
  uint64_t* q_full = bars;                     // 1
  uint64_t* k_full = bars + 1;                 // KS
  uint64_t* k_empty = k_full + KS;
  uint64_t* v_full = k_empty + KS;             // VS
  uint64_t* v_empty = v_full + VS;
  uint64_t* s_full = v_empty + VS;      // 2 (buffer == group == tile parity)
  uint64_t* s_empty = s_full + 2;              // 2
  uint64_t* p_full = s_empty + 2;              // 2
  uint64_t* p_empty = p_full + 2;              // 2  ("PV of this group's tile retired": P buffer free, O_g stable)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(p_empty + 2);

  const int warp = uniform_warp_index(), lane = threadIdx.x & 31;
  // Frames are taken in reverse launch order when a per-frame second segment exists: in the CFG batch the frames without it
  // (half the key tiles) come first, and CTAs are dispatched in grid order, so the long CTAs would otherwise form the tail.
  const int n = (args.has_seg2 && args.seg2_index) ? (int)(gridDim.z - 1 - blockIdx.z) : (int)blockIdx.z;
  const int h = blockIdx.y, q0 = blockIdx.x * BQ;

  if (threadIdx.x == 0) {
    mbar_init(q_full, 1);
    for (int s = 0; s < KS; ++s) { mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1); }
    for (int s = 0; s < VS; ++s) { mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 1); }
    for (int g = 0; g < 2; ++g) {
      mbar_init(&s_full[g], 1); mbar_init(&s_empty[g], 4); mbar_init(&p_full[g], 4); mbar_init(&p_empty[g], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_prologue();   // barriers and tensor memory are set up; the first global access (seg2_index, TMA) is below

  int seg2 = -1;
  if (args.has_seg2) seg2 = args.seg2_index ? args.seg2_index[n] : 0;
  const int tiles1 = (args.Lk + BN - 1) / BN;
  const int tiles2 = seg2 >= 0 ? (args.Lk2 + BN - 1) / BN : 0;
  const int num_tiles = tiles1 + tiles2;

  if (warp == 0 && elect_one_sync()) {
    // ===================== TMA producer =====================
    mbar_expect_tx(q_full, Q_BYTES);
#pragma unroll
    for (int c = 0; c < DCH; ++c) tma_load_3d(sQ + c * (BQ * 128), &tmQ, q_full, c * 64, h, n * args.Lq + q0);
    // Both rings are kept full: K(j + KS) is requested as soon as QK(j) retires, V(j + VS) as soon as PV(j) does.
    auto tile_row = [&](int j) {
      const bool second = j >= tiles1;
      return second ? seg2 * args.Lk2 + (j - tiles1) * BN : n * args.Lk + j * BN;
    };
    auto load_k = [&](int j) {
      const int st = j % KS;
      const bool second = j >= tiles1;
      mbar_wait(&k_empty[st], ((j / KS) & 1) ^ 1);
      mbar_expect_tx(&k_full[st], KV_BYTES);
#pragma unroll
      for (int c = 0; c < DCH; ++c)
        tma_load_3d(sK + st * KV_BYTES + c * KV_CHUNK, second ? &tmK2 : &tmK, &k_full[st], c * 64, h, tile_row(j));
    };
    auto load_v = [&](int j) {
      const int st = j % VS;
      const bool second = j >= tiles1;
      mbar_wait(&v_empty[st], ((j / VS) & 1) ^ 1);
      mbar_expect_tx(&v_full[st], KV_BYTES);
#pragma unroll
      for (int c = 0; c < DCH; ++c)
        tma_load_3d(sV + st * KV_BYTES + c * KV_CHUNK, second ? &tmV2 : &tmV, &v_full[st], c * 64, h, tile_row(j));
    };
    // The MMA warp retires QK(2), PV(0), QK(3), PV(1), ...: K(i + 2 + KS) can be requested once QK(i + 2) is done and
    // V(i + VS) once PV(i) is; requesting in that order keeps either ring from waiting behind the other.
    for (int j = 0; j < KS && j < num_tiles; ++j) load_k(j);
    for (int j = 0; j < VS && j < num_tiles; ++j) load_v(j);
    for (int j = KS; j < KS + 2 && j < num_tiles; ++j) load_k(j);
    for (int i = 0; i < num_tiles; ++i) {
      if (i + VS < num_tiles) load_v(i + VS);
      if (i + 2 + KS < num_tiles) load_k(i + 2 + KS);
    }
  } else if (warp == 1 && elect_one_sync()) {
    // ===================== MMA issuer =====================
    // Issue order: QK(0), QK(1), then per tile j: QK(j+2), PV(j).  QK(j+2) only needs group (j & 1) to have pulled
    // S(j) out of TMEM (signalled three quarters into its softmax of tile j), so S(j+2) is ready when the group
    // finishes tile j.  (An event-driven issue order, strict ping-pong of the two groups' exp phases, three / four
    // softmax groups on 64-key tiles and shared-memory P were all measured: 2.0-3.0 ms vs 2.0 ms for this form.)
    auto issue_qk = [&](int j) {
      const int st = j % KS, g = j & 1;
      tc_fence_after();
      const uint32_t tS = tmem_base + TM_S + g * BN;
      // descriptors differ only in the (address >> 4) field: build one per operand and step it
      const uint64_t dQ = desc_kmajor(smem_u32(sQ)), dK = desc_kmajor(smem_u32(sK + st * KV_BYTES));
#pragma unroll
      for (int k = 0; k < DPAD / 16; ++k) {
        if (k < qk_steps) {
          const uint32_t offq = (k / 4) * (BQ * 128) + (k % 4) * 32, offk = (k / 4) * KV_CHUNK + (k % 4) * 32;
          umma(tS, dQ + (offq >> 4), dK + (offk >> 4), IDESC_QK, k != 0);
        }
      }
      umma_commit(&k_empty[st]);
      umma_commit(&s_full[g]);
    };
    auto issue_pv = [&](int j) {
      const int st = j % VS, g = j & 1;
      tc_fence_after();
      const uint64_t dV = desc_mnmajor(smem_u32(sV + st * KV_BYTES), KV_CHUNK);
      const uint32_t tOg = tmem_base + TM_O + g * DPAD;
      if constexpr (P_TMEM) {
        const uint32_t tPg = tmem_base + TM_P + g * (BN / 2);
#pragma unroll
        for (int k = 0; k < BN / 16; ++k)   // A = P from TMEM: 16 keys = 8 packed columns per step; B = V, MN-major
          umma_ts(tOg, tPg + k * 8, dV + k * (2048 >> 4), IDESC_PV, ((j >> 1) | k) != 0);
      } else {
        const uint64_t dP = desc_kmajor(smem_u32(sP + g * P_BYTES));
#pragma unroll
        for (int k = 0; k < BN / 16; ++k) {
          // A = P: K-major, 64-key chunks of (128 rows x 128 B); B = V: MN-major, 16 keys = 2 x 1024 B per step
          const uint32_t offp = (k / 4) * (BQ * 128) + (k % 4) * 32;
          umma(tOg, dP + (offp >> 4), dV + k * (2048 >> 4), IDESC_PV, ((j >> 1) | k) != 0);
        }
      }
      umma_commit(&v_empty[st]);
      umma_commit(&p_empty[g]);
    };
    mbar_wait(q_full, 0);
    auto wait_qk = [&](int j) {
      mbar_wait(&k_full[j % KS], (j / KS) & 1);
      mbar_wait(&s_empty[j & 1], ((j >> 1) & 1) ^ 1);
      issue_qk(j);
    };
    wait_qk(0);
    if (num_tiles > 1) wait_qk(1);
    for (int j = 0; j < num_tiles; ++j) {
      if (j + 2 < num_tiles) wait_qk(j + 2);
      mbar_wait(&v_full[j % VS], (j / VS) & 1);
      mbar_wait(&p_full[j & 1], (j >> 1) & 1);
      issue_pv(j);
    }
  } else if (warp >= 2) {
    // ===================== softmax groups / merge / epilogue =====================
    const int g = (warp - 2) >> 2;                   // softmax group: tiles j = g, g+2, ...
    const int lane_grp = warp & 3;
    const int row = lane_grp * 32 + lane;            // query row inside the tile == TMEM lane
    const uint32_t lane_addr = (uint32_t)(lane_grp * 32) << 16;
    const float c = args.scale_log2e;
    float m_ref = -INFINITY;   // reference maximum the stored O_g / l are relative to (raw score units)
    float l_run = 0.f;
    const uint32_t prow_s = smem_u32(sP + g * P_BYTES + row * 128);
    const int sw = row & 7;
    const uint32_t tO = tmem_base + TM_O + g * DPAD + lane_addr;
    int it = 0;

    for (int j = g; j < num_tiles; j += 2, ++it) {
      const bool second = j >= tiles1;
      const int seg_len = second ? args.Lk2 : args.Lk;
      const int k0 = (second ? (j - tiles1) : j) * BN;
      const int valid = min(BN, seg_len - k0);       // keys of this tile that exist
      mbar_wait(&s_full[g], it & 1);
      tc_fence_after();
      // Online softmax in 32-column chunks, software-pipelined against the TMEM reads: tcgen05.ld moves only
      // 16 B/clk per scheduler (a 128 x 128 fp32 tile costs ~1000 clk, as much as its MUFU.EX2 work), so the load
      // of chunk i+1 is in flight while chunk i goes through max / exp2 / pack.  Each chunk is scored against the
      // running reference m_ref; when a chunk maximum exceeds it by more than 2^THRESHOLD (first tile, then rare)
      // the reference moves and what this tile already produced (packed P chunks, l) is rescaled in registers;
      // alpha_tile carries the factor the O accumulator in TMEM still owes.
      const uint32_t tS = tmem_base + TM_S + g * BN + lane_addr;
      uint32_t sa[32], sb[32];
      uint32_t pk[BN / 2];
      float alpha_tile = 1.f;
      float lsum[4] = {0.f, 0.f, 0.f, 0.f};
      auto chunk = [&](uint32_t (&sc)[32], const int ci) {
        if (valid < BN) {   // partial last tile of a segment: mask the keys that do not exist (warp-uniform branch)
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (ci * 32 + i >= valid) sc[i] = 0xff800000u;   // -inf
        }
        float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
#pragma unroll
          for (int e = 0; e < 4; ++e) mx4[e] = fmaxf(mx4[e], __uint_as_float(sc[i + e]));
        }
        const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
        if ((mx - m_ref) * c > RESCALE_THRESHOLD) {              // true for the first chunk ever (m_ref = -inf)
          const float a = ex2_approx((m_ref - mx) * c);          // 0 the first time
          m_ref = mx;
          alpha_tile *= a;
          l_run *= a;
#pragma unroll
          for (int e = 0; e < 4; ++e) lsum[e] *= a;
          const __nv_bfloat162 a2 = __float2bfloat162_rn(a);
#pragma unroll
          for (int i = 0; i < BN / 2; ++i) {
            if (i < ci * 16) {                                    // chunks of this tile that are already packed
              __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&pk[i]);
              v = __hmul2(v, a2);
              pk[i] = *reinterpret_cast<uint32_t*>(&v);
            }
          }
        }
        const float mc = m_ref * c;
#pragma unroll
        for (int i = 0; i < 32; i += 4) exp4_pack_sum<PACKED>(sc + i, c, mc, pk + ci * 16 + i / 2, lsum);
      };
      tmem_ld32(tS, sa);
      tmem_ld_wait();
      pin32(sa);
#pragma unroll
      for (int ci = 0; ci < BN / 32; ci += 2) {
        tmem_ld32(tS + (ci + 1) * 32, sb);                        // in flight during chunk ci
        chunk(sa, ci);
        tmem_ld_wait();
        pin32(sb);
        if (ci + 2 < BN / 32) {
          tmem_ld32(tS + (ci + 2) * 32, sa);                      // in flight during chunk ci + 1
        } else {
          tc_fence_before();                                      // S is in registers: hand the buffer back
          __syncwarp();
          if (lane == 0) mbar_arrive(&s_empty[g]);
        }
        chunk(sb, ci + 1);
        if (ci + 2 < BN / 32) {
          tmem_ld_wait();
          pin32(sa);
        }
      }
      l_run += (lsum[0] + lsum[1]) + (lsum[2] + lsum[3]);

      if (it > 0) {
        mbar_wait(&p_empty[g], (it - 1) & 1);   // this group's previous PV retired: P buffer reusable, O_g stable
        tc_fence_after();
        if (__any_sync(0xffffffffu, alpha_tile != 1.f)) {
#pragma unroll
          for (int cc = 0; cc < DPAD / 32; ++cc) {
            uint32_t o[32];
            tmem_ld32(tO + cc * 32, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha_tile);
            tmem_st32(tO + cc * 32, o);
          }
          tmem_st_wait();
        }
      }
      if constexpr (P_TMEM) {
        // P -> TMEM: lane = query row, column = key pair
        const uint32_t tP = tmem_base + TM_P + g * (BN / 2) + lane_addr;
#pragma unroll
        for (int i = 0; i < BN / 64; ++i) tmem_st32(tP + i * 32, pk + i * 32);
        tmem_st_wait();
      } else {
        // P -> smem, K-major 128B swizzle: 16-byte unit u of row r lands at unit (u ^ (r & 7)) of its 128-byte row
#pragma unroll
        for (int u = 0; u < BN / 8; ++u) {
          const int chunk_id = u >> 3, uu = u & 7;
          st_shared_v4(prow_s + chunk_id * (BQ * 128) + ((uu ^ sw) << 4), pk[u * 4], pk[u * 4 + 1], pk[u * 4 + 2], pk[u * 4 + 3]);
        }
        fence_async_smem();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[g]);
    }

    // ---- merge the two groups and store O / l (only the first d columns are real)
    if (it > 0) {
      mbar_wait(&p_empty[g], (it - 1) & 1);
      tc_fence_after();
    }
    if (g == 1) {
      stats[2 * row] = m_ref;
      stats[2 * row + 1] = l_run;
    }
    tc_fence_before();
    asm volatile("bar.sync 1, 256;" ::: "memory");     // the 8 softmax warps
    tc_fence_after();
    if (g == 0) {
      const float m1 = stats[2 * row], l1 = stats[2 * row + 1];
      const bool has1 = num_tiles > 1;                 // group 1 processed at least one tile (uniform)
      const float m = has1 ? fmaxf(m_ref, m1) : m_ref;
      const float a0 = ex2_approx((m_ref - m) * c);
      const float a1 = has1 ? ex2_approx((m1 - m) * c) : 0.f;
      const float inv = 1.f / (l_run * a0 + (has1 ? l1 * a1 : 0.f));
      const float w0 = a0 * inv, w1 = a1 * inv;
      const int q = q0 + row;
      bf16* orow = args.out + ((int64_t)n * args.Lq + q) * args.ldo + h * args.d;
      const uint32_t tO0 = tmem_base + TM_O + lane_addr, tO1 = tO0 + DPAD;
#pragma unroll
      for (int cc = 0; cc < DPAD / 32; ++cc) {
        uint32_t o[32], o1[32];
        tmem_ld32(tO0 + cc * 32, o);
        if (has1) tmem_ld32(tO1 + cc * 32, o1);
        tmem_ld_wait();
        float f[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(o[i]) * w0 + (has1 ? __uint_as_float(o1[i]) * w1 : 0.f);
        if (q < args.Lq) {
#pragma unroll
          for (int gg = 0; gg < 4; ++gg) {
            const int col = cc * 32 + gg * 8;
            if (col < args.d) {
              uint4 val;
              val.x = pack_bf16(f[gg * 8 + 0], f[gg * 8 + 1]);
              val.y = pack_bf16(f[gg * 8 + 2], f[gg * 8 + 3]);
              val.z = pack_bf16(f[gg * 8 + 4], f[gg * 8 + 5]);
              val.w = pack_bf16(f[gg * 8 + 6], f[gg * 8 + 7]);
              *reinterpret_cast<uint4*>(orow + col) = val;
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

// ------------------------------------------------------------------------------------------- head dim <= 64, version 2
// Same contract as attention_tc_kernel<1, 128>, restructured after the ncu source-level profile of that kernel at the
// config-2 level-0 shape (profiles/r2_attention_v2.md): 23 % of all warp samples sat in the softmax groups' wait for
// S = Q K^T.  With one S buffer per group QK(j+2) could only be issued once the group had pulled S(j) into registers
// (three quarters into its tile), leaving ~1/4 tile of slack against the ~450-clock mbarrier -> MMA -> commit -> wake-up
// latency.  Here S lives in THREE rotating 128-column buffers and P (bf16, 64 columns) is written back over the S buffer
// it was computed from (the TMEM budget is exactly 3 x 128 + 2 x 64 = 512 columns): QK(j+3) is issued as soon as
// PV(j) has retired, i.e. a HALF tile period before its softmax group asks for it.
// The running-max bookkeeping is also simpler (smaller code: instruction-fetch stalls were 7 % of the samples): chunks
// are scored against the reference maximum m_ref; when a chunk exceeds it by more than 2^8 (always on the first tile,
// rarely later) the thread takes the exact row maximum of the tile from a second pass over S -- still in TMEM, because
// the buffer is only released by writing P -- rescales l and the pending factor for O, and redoes the tile once.
template <int BN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
attention_tc64_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                      const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmK2,
                      const __grid_constant__ CUtensorMap tmV2, const AttnArgs args) {
  static_assert(BN == 128, "three 128-column S buffers + two 64-column O accumulators fill the 512 TMEM columns");
  constexpr int DPAD = 64, KS = 4, VS = 3, NB = 3;
  constexpr int Q_BYTES = BQ * DPAD * 2;
  constexpr int KV_BYTES = BN * DPAD * 2;
  constexpr int TM_O = NB * BN;                   // O accumulators at 384 / 448
  constexpr uint32_t IDESC_QK = idesc_bf16(BN, false);
  const int qk_steps = (args.d + 15) >> 4;
  const uint32_t IDESC_PV = idesc_bf16(qk_steps << 4, true);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + Q_BYTES;
  uint8_t* sV = sK + KS * KV_BYTES;
  float* stats = reinterpret_cast<float*>(sV + VS * KV_BYTES);        // [128][2]: (m_ref, l) of group 1
  uint64_t* bars = reinterpret_cast<uint64_t*>(stats + 2 * BQ);
  uint64_t* q_full = bars;                     // 1
  uint64_t* k_full = bars + 1;                 // KS
  uint64_t* k_empty = k_full + KS;
  uint64_t* v_full = k_empty + KS;             // VS
  uint64_t* v_empty = v_full + VS;
  uint64_t* s_full = v_empty + VS;             // NB: S(j) complete in buffer j % NB
  uint64_t* p_full = s_full + NB;              // NB: P(j) written over it (4 warps arrive)
  uint64_t* o_ready = p_full + NB;             // 2:  PV of the group's latest tile retired (O_g stable, buffer reusable)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_ready + 2);

  const int warp = uniform_warp_index(), lane = threadIdx.x & 31;
  // Frames are taken in reverse launch order when a per-frame second segment exists: in the CFG batch the frames without it
  // (half the key tiles) come first, and CTAs are dispatched in grid order, so the long CTAs would otherwise form the tail.
  const int n = (args.has_seg2 && args.seg2_index) ? (int)(gridDim.z - 1 - blockIdx.z) : (int)blockIdx.z;
  const int h = blockIdx.y, q0 = blockIdx.x * BQ;

  if (threadIdx.x == 0) {
    mbar_init(q_full, 1);
    for (int s = 0; s < KS; ++s) { mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1); }
    for (int s = 0; s < VS; ++s) { mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 1); }
    for (int b = 0; b < NB; ++b) { mbar_init(&s_full[b], 1); mbar_init(&p_full[b], 4); }
    mbar_init(&o_ready[0], 1); mbar_init(&o_ready[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_prologue();

  int seg2 = -1;
  if (args.has_seg2) seg2 = args.seg2_index ? args.seg2_index[n] : 0;
  const int tiles1 = (args.Lk + BN - 1) / BN;
  const int tiles2 = seg2 >= 0 ? (args.Lk2 + BN - 1) / BN : 0;
  const int num_tiles = tiles1 + tiles2;

  if (warp == 0 && elect_one_sync()) {
    // ===================== TMA producer =====================
    mbar_expect_tx(q_full, Q_BYTES);
    tma_load_3d(sQ, &tmQ, q_full, 0, h, n * args.Lq + q0);
    auto tile_row = [&](int j) { return j >= tiles1 ? seg2 * args.Lk2 + (j - tiles1) * BN : n * args.Lk + j * BN; };
    auto load_k = [&](int j) {
      const int st = j % KS;
      mbar_wait(&k_empty[st], ((j / KS) & 1) ^ 1);
      mbar_expect_tx(&k_full[st], KV_BYTES);
      tma_load_3d(sK + st * KV_BYTES, j >= tiles1 ? &tmK2 : &tmK, &k_full[st], 0, h, tile_row(j));
    };
    auto load_v = [&](int j) {
      const int st = j % VS;
      mbar_wait(&v_empty[st], ((j / VS) & 1) ^ 1);
      mbar_expect_tx(&v_full[st], KV_BYTES);
      tma_load_3d(sV + st * KV_BYTES, j >= tiles1 ? &tmV2 : &tmV, &v_full[st], 0, h, tile_row(j));
    };
    // The MMA warp consumes K(0..2), then per tile j: V(j), K(j+3).  K(i+4) can be requested once QK(i) retired (early),
    // V(i+3) once PV(i) did.
    for (int j = 0; j < KS && j < num_tiles; ++j) load_k(j);
    for (int j = 0; j < VS && j < num_tiles; ++j) load_v(j);
    for (int i = 0; i < num_tiles; ++i) {
      if (i + KS < num_tiles) load_k(i + KS);
      if (i + VS < num_tiles) load_v(i + VS);
    }
  } else if (warp == 1 && elect_one_sync()) {
    // ===================== MMA issuer =====================
    auto issue_qk = [&](int j) {
      const int st = j % KS, b = j % NB;
      mbar_wait(&k_full[st], (j / KS) & 1);
      tc_fence_after();
      const uint32_t tS = tmem_base + b * BN;
      const uint64_t dQ = desc_kmajor(smem_u32(sQ)), dK = desc_kmajor(smem_u32(sK + st * KV_BYTES));
#pragma unroll
      for (int k = 0; k < DPAD / 16; ++k)
        if (k < qk_steps) umma(tS, dQ + 2 * k, dK + 2 * k, IDESC_QK, k != 0);
      umma_commit(&k_empty[st]);
      umma_commit(&s_full[b]);
    };
    mbar_wait(q_full, 0);
    for (int j = 0; j < NB && j < num_tiles; ++j) issue_qk(j);
    for (int j = 0; j < num_tiles; ++j) {
      const int st = j % VS, b = j % NB, g = j & 1;
      mbar_wait(&v_full[st], (j / VS) & 1);
      mbar_wait(&p_full[b], (j / NB) & 1);
      tc_fence_after();
      const uint64_t dV = desc_mnmajor(smem_u32(sV + st * KV_BYTES), BN * 128);
      const uint32_t tOg = tmem_base + TM_O + g * DPAD, tP = tmem_base + b * BN;
#pragma unroll
      for (int k = 0; k < BN / 16; ++k)   // A = P from TMEM: 16 keys = 8 packed columns per step; B = V, MN-major
        umma_ts(tOg, tP + k * 8, dV + k * (2048 >> 4), IDESC_PV, ((j >> 1) | k) != 0);
      umma_commit(&v_empty[st]);
      umma_commit(&o_ready[g]);
      if (j + NB < num_tiles) {
        // QK(j + 3) overwrites the buffer PV(j) reads P from: issue it once PV(j) has retired (half a tile period before
        // its softmax group needs it)
        mbar_wait(&o_ready[g], (j >> 1) & 1);
        issue_qk(j + NB);
      }
    }
  } else if (warp >= 2) {
    // ===================== softmax groups / merge / epilogue =====================
    const int g = (warp - 2) >> 2;                   // softmax group: tiles j = g, g+2, ...
    const int lane_grp = warp & 3;
    const int row = lane_grp * 32 + lane;            // query row inside the tile == TMEM lane
    const uint32_t lane_addr = (uint32_t)(lane_grp * 32) << 16;
    const float c = args.scale_log2e;
    float m_ref = -INFINITY;   // reference maximum the stored O_g / l are relative to (raw score units)
    float l_run = 0.f;
    const uint32_t tO = tmem_base + TM_O + g * DPAD + lane_addr;
    int it = 0;

    for (int j = g; j < num_tiles; j += 2, ++it) {
      const bool second = j >= tiles1;
      const int seg_len = second ? args.Lk2 : args.Lk;
      const int k0 = (second ? (j - tiles1) : j) * BN;
      const int valid = min(BN, seg_len - k0);       // keys of this tile that exist
      const int b = j % NB;
      const uint32_t tS = tmem_base + b * BN + lane_addr;
      mbar_wait(&s_full[b], (j / NB) & 1);
      tc_fence_after();
      uint32_t pk[BN / 2];
      float alpha_tile = 1.f;
      float lsum[4];
      bool exceeded = false;
      auto process = [&]() {
        lsum[0] = lsum[1] = lsum[2] = lsum[3] = 0.f;
        const float mc = m_ref * c;
        uint32_t sa[32], sb[32];
        auto chunk = [&](uint32_t (&sc)[32], const int ci) {
          if (valid < BN) {   // partial last tile of a segment: mask the keys that do not exist (warp-uniform branch)
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (ci * 32 + i >= valid) sc[i] = 0xff800000u;   // -inf
          }
          float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
#pragma unroll
            for (int e = 0; e < 4; ++e) mx4[e] = fmaxf(mx4[e], __uint_as_float(sc[i + e]));
          }
          const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
          exceeded = exceeded || ((mx - m_ref) * c > RESCALE_THRESHOLD);   // true for the first tile (m_ref = -inf)
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float p0 = ex2_approx(fmaf(__uint_as_float(sc[i]), c, -mc));
            const float p1 = ex2_approx(fmaf(__uint_as_float(sc[i + 1]), c, -mc));
            const float p2 = ex2_approx(fmaf(__uint_as_float(sc[i + 2]), c, -mc));
            const float p3 = ex2_approx(fmaf(__uint_as_float(sc[i + 3]), c, -mc));
            pk[ci * 16 + i / 2] = pack_bf16(p0, p1);
            pk[ci * 16 + i / 2 + 1] = pack_bf16(p2, p3);
            lsum[0] += p0; lsum[1] += p1; lsum[2] += p2; lsum[3] += p3;
          }
        };
        tmem_ld32(tS, sa);
        tmem_ld_wait();
        pin32(sa);
#pragma unroll
        for (int ci = 0; ci < BN / 32; ci += 2) {
          tmem_ld32(tS + (ci + 1) * 32, sb);                        // in flight during chunk ci
          chunk(sa, ci);
          tmem_ld_wait();
          pin32(sb);
          if (ci + 2 < BN / 32) tmem_ld32(tS + (ci + 2) * 32, sa);  // in flight during chunk ci + 1
          chunk(sb, ci + 1);
          if (ci + 2 < BN / 32) {
            tmem_ld_wait();
            pin32(sa);
          }
        }
      };
      process();
      if (__any_sync(0xffffffffu, exceeded)) {
        // Rare after the first tile.  Exact row maximum of this tile from a second pass over S (still in TMEM), then the
        // tile is redone against it; what the previous tiles accumulated (l, O) is rescaled by 2^((m_old - m_new) c).
        float mx = -INFINITY;
#pragma unroll 1
        for (int ci = 0; ci < BN / 32; ++ci) {
          uint32_t sa[32];
          tmem_ld32(tS + ci * 32, sa);
          tmem_ld_wait();
          pin32(sa);
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (ci * 32 + i < valid) mx = fmaxf(mx, __uint_as_float(sa[i]));
        }
        if ((mx - m_ref) * c > RESCALE_THRESHOLD) {      // (lanes whose row did not move keep their reference)
          const float a = ex2_approx((m_ref - mx) * c);  // 0 the first time (m_ref = -inf)
          m_ref = mx;
          alpha_tile *= a;
          l_run *= a;
        }
        process();                                       // every lane redoes the tile (warp-uniform TMEM loads)
      }
      l_run += (lsum[0] + lsum[1]) + (lsum[2] + lsum[3]);

      if (it > 0) {
        mbar_wait(&o_ready[g], (it - 1) & 1);   // this group's previous PV retired: O_g stable
        tc_fence_after();
        if (__any_sync(0xffffffffu, alpha_tile != 1.f)) {
#pragma unroll
          for (int cc = 0; cc < DPAD / 32; ++cc) {
            uint32_t o[32];
            tmem_ld32(tO + cc * 32, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha_tile);
            tmem_st32(tO + cc * 32, o);
          }
          tmem_st_wait();
        }
      }
      // P over the S buffer it came from: lane = query row, column = key pair
#pragma unroll
      for (int i = 0; i < BN / 64; ++i) tmem_st32(tS + i * 32, pk + i * 32);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[b]);
    }

    // ---- merge the two groups and store O / l (only the first d columns are real)
    if (it > 0) {
      mbar_wait(&o_ready[g], (it - 1) & 1);
      tc_fence_after();
    }
    if (g == 1) {
      stats[2 * row] = m_ref;
      stats[2 * row + 1] = l_run;
    }
    tc_fence_before();
    asm volatile("bar.sync 1, 256;" ::: "memory");     // the 8 softmax warps
    tc_fence_after();
    if (g == 0) {
      const float m1 = stats[2 * row], l1 = stats[2 * row + 1];
      const bool has1 = num_tiles > 1;                 // group 1 processed at least one tile (uniform)
      const float m = has1 ? fmaxf(m_ref, m1) : m_ref;
      const float a0 = ex2_approx((m_ref - m) * c);
      const float a1 = has1 ? ex2_approx((m1 - m) * c) : 0.f;
      const float inv = 1.f / (l_run * a0 + (has1 ? l1 * a1 : 0.f));
      const float w0 = a0 * inv, w1 = a1 * inv;
      const int q = q0 + row;
      bf16* orow = args.out + ((int64_t)n * args.Lq + q) * args.ldo + h * args.d;
      const uint32_t tO0 = tmem_base + TM_O + lane_addr, tO1 = tO0 + DPAD;
#pragma unroll
      for (int cc = 0; cc < DPAD / 32; ++cc) {
        if (cc * 32 < args.d) {
          uint32_t o[32], o1[32];
          tmem_ld32(tO0 + cc * 32, o);
          if (has1) tmem_ld32(tO1 + cc * 32, o1);
          tmem_ld_wait();
          float f[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(o[i]) * w0 + (has1 ? __uint_as_float(o1[i]) * w1 : 0.f);
          if (q < args.Lq) {
#pragma unroll
            for (int gg = 0; gg < 4; ++gg) {
              const int col = cc * 32 + gg * 8;
              if (col < args.d) {
                uint4 val;
                val.x = pack_bf16(f[gg * 8 + 0], f[gg * 8 + 1]);
                val.y = pack_bf16(f[gg * 8 + 2], f[gg * 8 + 3]);
                val.z = pack_bf16(f[gg * 8 + 4], f[gg * 8 + 5]);
                val.w = pack_bf16(f[gg * 8 + 6], f[gg * 8 + 7]);
                *reinterpret_cast<uint4*>(orow + col) = val;
              }
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

// ------------------------------------------------------------------------------------- head dim <= 64, persistent CTAs
// Same contract and the same per-tile pipeline as attention_tc_kernel<1, 128>, but ONE CTA per SM that walks a list of
// (frame, head, 128-query tile) items.  The ncu source-level profile of the one-item-per-CTA kernel at the config-2 level-0
// shape (profiles/r2_ops_ncu.md, 6144 CTAs of 22-44 us on 148 SMs) shows ~5 % of all warp samples on the CTA-end
// barrier / EXIT and another ~4 % on the wait for the first S of a CTA: with a single resident CTA per SM (512 TMEM
// columns) nothing overlaps the barrier / TMEM set-up, the Q and first K loads and the merge-and-store tail of
// consecutive CTAs.  Here the rings, barriers and TMEM live for the whole launch:
//   * the producer runs ahead into the next item (Q is double buffered; the K / V rings simply continue), so the next
//     item's first tiles are in shared memory before the current item's last PV retires;
//   * the MMA warp issues QK(0), QK(1) of the next item as soon as the softmax groups have pulled the last S tiles of the
//     current one, i.e. while group 0 still merges and stores;
//   * group 1 starts the next item's tile 1 right after handing its (m, l) to group 0, so the MUFU pipe stays fed during
//     the tail.
// Ordering of the O accumulators across items needs no extra barrier: the first MMA that touches them again is PV(0) of
// the next item (PV(1) is issued after it, the MMA warp issues in order), which waits for P(0) of that item, which
// group 0 writes only after its merge has read O_0 and O_1 (tcgen05.wait::ld, then fence + mbarrier arrive).
// Barrier phases run over launch-wide counters: K / V tiles (ring stages), tiles per softmax group (S / P hand-offs),
// items (Q buffers, statistics buffers).  Items are dealt round-robin, the frames with a second segment first.
template <int BN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
attention_tc_persist_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                            const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmK2,
                            const __grid_constant__ CUtensorMap tmV2, const AttnArgs args, const int q_tiles,
                            const int num_items) {
  static_assert(BN == 128, "2 S buffers + 2 O accumulators + 2 packed P buffers fill the 512 TMEM columns");
  constexpr int DPAD = 64, KS = 4, VS = 3;
  constexpr int Q_BYTES = BQ * DPAD * 2;
  constexpr int KV_BYTES = BN * DPAD * 2;
  constexpr int TM_S = 0, TM_O = 2 * BN, TM_P = TM_O + 2 * DPAD;
  constexpr uint32_t IDESC_QK = idesc_bf16(BN, false);
  const int qk_steps = (args.d + 15) >> 4;
  const uint32_t IDESC_PV = idesc_bf16(qk_steps << 4, true);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem;                                                  // 2 buffers (item parity)
  uint8_t* sK = sQ + 2 * Q_BYTES;
  uint8_t* sV = sK + KS * KV_BYTES;
  float* stats = reinterpret_cast<float*>(sV + VS * KV_BYTES);        // [item parity][128][2]: (m_ref, l) of group 1
  uint64_t* bars = reinterpret_cast<uint64_t*>(stats + 2 * 2 * BQ);
  uint64_t* q_full = bars;                     // 2
  uint64_t* q_empty = q_full + 2;              // 2  (all QK of the item retired)
  uint64_t* k_full = q_empty + 2;              // KS
  uint64_t* k_empty = k_full + KS;
  uint64_t* v_full = k_empty + KS;             // VS
  uint64_t* v_empty = v_full + VS;
  uint64_t* s_full = v_empty + VS;             // 2 (buffer == group == tile parity inside the item)
  uint64_t* s_empty = s_full + 2;
  uint64_t* p_full = s_empty + 2;
  uint64_t* p_empty = p_full + 2;              // 2  ("PV of this group's tile retired": P buffer free, O_g stable)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(p_empty + 2);

  const int warp = uniform_warp_index(), lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int b = 0; b < 2; ++b) { mbar_init(&q_full[b], 1); mbar_init(&q_empty[b], 1); }
    for (int s = 0; s < KS; ++s) { mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1); }
    for (int s = 0; s < VS; ++s) { mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 1); }
    for (int g = 0; g < 2; ++g) {
      mbar_init(&s_full[g], 1); mbar_init(&s_empty[g], 4); mbar_init(&p_full[g], 4); mbar_init(&p_empty[g], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_prologue();

  // item -> (frame, head, first query row, second-segment row or -1, key tiles per segment); same for every role
  const int per_frame = args.heads * q_tiles;
  const bool reversed = args.has_seg2 && args.seg2_index;             // CFG batch: the frames without a second segment come first
  struct Item { int n, h, q0, seg2, tiles1, tiles2; };
  auto decode = [&](int w) {
    Item it;
    const int z = w / per_frame, r = w - z * per_frame;
    it.n = reversed ? args.N - 1 - z : z;
    it.h = r / q_tiles;
    it.q0 = (r - it.h * q_tiles) * BQ;
    it.seg2 = -1;
    if (args.has_seg2) it.seg2 = args.seg2_index ? __ldg(args.seg2_index + it.n) : 0;
    it.tiles1 = (args.Lk + BN - 1) / BN;
    it.tiles2 = it.seg2 >= 0 ? (args.Lk2 + BN - 1) / BN : 0;
    return it;
  };

  if (warp == 0 && elect_one_sync()) {
    // ===================== TMA producer =====================
    uint32_t kt0 = 0, vt0 = 0;                 // K / V tiles requested before this item
    int itn = 0;
    for (int w = blockIdx.x; w < num_items; w += gridDim.x, ++itn) {
      const Item im = decode(w);
      const int num_tiles = im.tiles1 + im.tiles2;
      const int qb = itn & 1;
      mbar_wait(&q_empty[qb], ((itn >> 1) & 1) ^ 1);
      mbar_expect_tx(&q_full[qb], Q_BYTES);
      tma_load_3d(sQ + qb * Q_BYTES, &tmQ, &q_full[qb], 0, im.h, im.n * args.Lq + im.q0);
      auto tile_row = [&](int j) {
        const bool second = j >= im.tiles1;
        return second ? im.seg2 * args.Lk2 + (j - im.tiles1) * BN : im.n * args.Lk + j * BN;
      };
      auto load_k = [&](int j) {
        const uint32_t t = kt0 + j, st = t % KS;
        mbar_wait(&k_empty[st], ((t / KS) & 1) ^ 1);
        mbar_expect_tx(&k_full[st], KV_BYTES);
        tma_load_3d(sK + st * KV_BYTES, j >= im.tiles1 ? &tmK2 : &tmK, &k_full[st], 0, im.h, tile_row(j));
      };
      auto load_v = [&](int j) {
        const uint32_t t = vt0 + j, st = t % VS;
        mbar_wait(&v_empty[st], ((t / VS) & 1) ^ 1);
        mbar_expect_tx(&v_full[st], KV_BYTES);
        tma_load_3d(sV + st * KV_BYTES, j >= im.tiles1 ? &tmV2 : &tmV, &v_full[st], 0, im.h, tile_row(j));
      };
      // request order = consumption order of the MMA warp (QK runs two tiles ahead of PV), see attention_tc_kernel
      for (int j = 0; j < KS && j < num_tiles; ++j) load_k(j);
      for (int j = 0; j < VS && j < num_tiles; ++j) load_v(j);
      for (int j = KS; j < KS + 2 && j < num_tiles; ++j) load_k(j);
      for (int i = 0; i < num_tiles; ++i) {
        if (i + VS < num_tiles) load_v(i + VS);
        if (i + 2 + KS < num_tiles) load_k(i + 2 + KS);
      }
      kt0 += num_tiles;
      vt0 += num_tiles;
    }
  } else if (warp == 1 && elect_one_sync()) {
    // ===================== MMA issuer =====================
    uint32_t kt0 = 0, vt0 = 0;
    uint32_t gc_a = 0, gc_b = 0;               // tiles handed to softmax group 0 / 1 before this item
    int itn = 0;
    for (int w = blockIdx.x; w < num_items; w += gridDim.x, ++itn) {
      const Item im = decode(w);
      const int num_tiles = im.tiles1 + im.tiles2;
      const int qb = itn & 1;
      const uint64_t dQ = desc_kmajor(smem_u32(sQ + qb * Q_BYTES));
      auto issue_qk = [&](int j) {
        const uint32_t t = kt0 + j, st = t % KS;
        const int g = j & 1;
        const uint32_t c = (g ? gc_b : gc_a) + (j >> 1);
        mbar_wait(&k_full[st], (t / KS) & 1);
        mbar_wait(&s_empty[g], (c & 1) ^ 1);
        tc_fence_after();
        const uint32_t tS = tmem_base + TM_S + g * BN;
        const uint64_t dK = desc_kmajor(smem_u32(sK + st * KV_BYTES));
#pragma unroll
        for (int k = 0; k < DPAD / 16; ++k)
          if (k < qk_steps) umma(tS, dQ + ((k * 32) >> 4), dK + ((k * 32) >> 4), IDESC_QK, k != 0);
        umma_commit(&k_empty[st]);
        umma_commit(&s_full[g]);
        if (j == num_tiles - 1) umma_commit(&q_empty[qb]);      // every QK of the item has been issued: Q buffer free once they retire
      };
      auto issue_pv = [&](int j) {
        const uint32_t t = vt0 + j, st = t % VS;
        const int g = j & 1;
        const uint32_t c = (g ? gc_b : gc_a) + (j >> 1);
        mbar_wait(&v_full[st], (t / VS) & 1);
        mbar_wait(&p_full[g], c & 1);
        tc_fence_after();
        const uint64_t dV = desc_mnmajor(smem_u32(sV + st * KV_BYTES), BN * 128);
        const uint32_t tOg = tmem_base + TM_O + g * DPAD, tPg = tmem_base + TM_P + g * (BN / 2);
#pragma unroll
        for (int k = 0; k < BN / 16; ++k)
          umma_ts(tOg, tPg + k * 8, dV + k * (2048 >> 4), IDESC_PV, ((j >> 1) | k) != 0);
        umma_commit(&v_empty[st]);
        umma_commit(&p_empty[g]);
      };
      mbar_wait(&q_full[qb], (itn >> 1) & 1);
      issue_qk(0);
      if (num_tiles > 1) issue_qk(1);
      for (int j = 0; j < num_tiles; ++j) {
        if (j + 2 < num_tiles) issue_qk(j + 2);
        issue_pv(j);
      }
      kt0 += num_tiles;
      vt0 += num_tiles;
      gc_a += (num_tiles + 1) >> 1;
      gc_b += num_tiles >> 1;
    }
  } else if (warp >= 2) {
    // ===================== softmax groups / merge / epilogue =====================
    const int g = (warp - 2) >> 2;
    const int lane_grp = warp & 3;
    const int row = lane_grp * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(lane_grp * 32) << 16;
    const float c = args.scale_log2e;
    // one live base register: O_g, P_g and S_g sit 64, 64 and 128 columns apart per group
    static_assert(DPAD == 64 && BN == 128, "group strides below");
    const uint32_t tG = tmem_base + lane_addr + g * 64;
#define tO (tG + TM_O)
#define tP (tG + TM_P)
#define tS (tG + g * 64 + TM_S)
    uint32_t cnt = 0;                          // tiles this group has taken since the launch began
    int itn = 0;
    for (int w = blockIdx.x; w < num_items; w += gridDim.x, ++itn) {
      int num_tiles;
      const int tiles1 = (args.Lk + BN - 1) / BN;
      {
        const Item im = decode(w);           // only the tile count stays live across the tile loop; re-decoded for the store
        num_tiles = im.tiles1 + im.tiles2;
      }
      float m_ref = -INFINITY, l_run = 0.f;
      int it = 0;
      for (int j = g; j < num_tiles; j += 2, ++it, ++cnt) {
        const bool second = j >= tiles1;
        const int seg_len = second ? args.Lk2 : args.Lk;
        const int k0 = (second ? (j - tiles1) : j) * BN;
        const int valid = min(BN, seg_len - k0);
        mbar_wait(&s_full[g], cnt & 1);
        tc_fence_after();
        uint32_t sa[32], sb[32];
        uint32_t pk[BN / 2];
        float alpha_tile = 1.f;
        float lsum[4] = {0.f, 0.f, 0.f, 0.f};
        auto chunk = [&](uint32_t (&sc)[32], const int ci) {
          if (valid < BN) {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (ci * 32 + i >= valid) sc[i] = 0xff800000u;   // -inf
          }
          float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
#pragma unroll
            for (int e = 0; e < 4; ++e) mx4[e] = fmaxf(mx4[e], __uint_as_float(sc[i + e]));
          }
          const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
          if ((mx - m_ref) * c > RESCALE_THRESHOLD) {
            const float a = ex2_approx((m_ref - mx) * c);
            m_ref = mx;
            alpha_tile *= a;
            l_run *= a;
#pragma unroll
            for (int e = 0; e < 4; ++e) lsum[e] *= a;
            const __nv_bfloat162 a2 = __float2bfloat162_rn(a);
#pragma unroll
            for (int i = 0; i < BN / 2; ++i) {
              if (i < ci * 16) {
                __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&pk[i]);
                v = __hmul2(v, a2);
                pk[i] = *reinterpret_cast<uint32_t*>(&v);
              }
            }
          }
          const float mc = m_ref * c;
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float p0 = ex2_approx(fmaf(__uint_as_float(sc[i]), c, -mc));
            const float p1 = ex2_approx(fmaf(__uint_as_float(sc[i + 1]), c, -mc));
            const float p2 = ex2_approx(fmaf(__uint_as_float(sc[i + 2]), c, -mc));
            const float p3 = ex2_approx(fmaf(__uint_as_float(sc[i + 3]), c, -mc));
            pk[ci * 16 + i / 2] = pack_bf16(p0, p1);
            pk[ci * 16 + i / 2 + 1] = pack_bf16(p2, p3);
            lsum[0] += p0; lsum[1] += p1; lsum[2] += p2; lsum[3] += p3;
          }
        };
        tmem_ld32(tS, sa);
        tmem_ld_wait();
        pin32(sa);
#pragma unroll
        for (int ci = 0; ci < BN / 32; ci += 2) {
          tmem_ld32(tS + (ci + 1) * 32, sb);
          chunk(sa, ci);
          tmem_ld_wait();
          pin32(sb);
          if (ci + 2 < BN / 32) {
            tmem_ld32(tS + (ci + 2) * 32, sa);
          } else {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_empty[g]);
          }
          chunk(sb, ci + 1);
          if (ci + 2 < BN / 32) {
            tmem_ld_wait();
            pin32(sa);
          }
        }
        l_run += (lsum[0] + lsum[1]) + (lsum[2] + lsum[3]);

        if (cnt > 0) {
          // this group's previous PV (of this item, or the last one of the previous item) retired: P buffer reusable
          mbar_wait(&p_empty[g], (cnt - 1) & 1);
          tc_fence_after();
          if (it > 0 && __any_sync(0xffffffffu, alpha_tile != 1.f)) {
#pragma unroll
            for (int cc = 0; cc < DPAD / 32; ++cc) {
              uint32_t o[32];
              tmem_ld32(tO + cc * 32, o);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha_tile);
              tmem_st32(tO + cc * 32, o);
            }
            tmem_st_wait();
          }
        }
#pragma unroll
        for (int i = 0; i < BN / 64; ++i) tmem_st32(tP + i * 32, pk + i * 32);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[g]);
      }

      // ---- merge the two groups and store O / l (only the first d columns are real)
      if (it > 0) {
        mbar_wait(&p_empty[g], (cnt - 1) & 1);
        tc_fence_after();
      }
      float* st_item = stats + (itn & 1) * 2 * BQ;
      if (g == 1) {
        st_item[2 * row] = m_ref;
        st_item[2 * row + 1] = l_run;
      }
      tc_fence_before();
      asm volatile("bar.sync 1, 256;" ::: "memory");
      tc_fence_after();
      if (g == 0) {
        const Item im = decode(w);
        const float m1 = st_item[2 * row], l1 = st_item[2 * row + 1];
        const bool has1 = num_tiles > 1;
        const float m = has1 ? fmaxf(m_ref, m1) : m_ref;
        const float a0 = ex2_approx((m_ref - m) * c);
        const float a1 = has1 ? ex2_approx((m1 - m) * c) : 0.f;
        const float inv = 1.f / (l_run * a0 + (has1 ? l1 * a1 : 0.f));
        const float w0 = a0 * inv, w1 = a1 * inv;
        const int q = im.q0 + row;
        bf16* orow = args.out + ((int64_t)im.n * args.Lq + q) * args.ldo + im.h * args.d;
        const uint32_t tO0 = tmem_base + TM_O + lane_addr, tO1 = tO0 + DPAD;
#pragma unroll
        for (int cc = 0; cc < DPAD / 32; ++cc) {
          if (cc * 32 < args.d) {
            uint32_t o[32], o1[32];
            tmem_ld32(tO0 + cc * 32, o);
            if (has1) tmem_ld32(tO1 + cc * 32, o1);
            tmem_ld_wait();
            float f[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(o[i]) * w0 + (has1 ? __uint_as_float(o1[i]) * w1 : 0.f);
            if (q < args.Lq) {
#pragma unroll
              for (int gg = 0; gg < 4; ++gg) {
                const int col = cc * 32 + gg * 8;
                if (col < args.d) {
                  uint4 val;
                  val.x = pack_bf16(f[gg * 8 + 0], f[gg * 8 + 1]);
                  val.y = pack_bf16(f[gg * 8 + 2], f[gg * 8 + 3]);
                  val.z = pack_bf16(f[gg * 8 + 4], f[gg * 8 + 5]);
                  val.w = pack_bf16(f[gg * 8 + 6], f[gg * 8 + 7]);
                  *reinterpret_cast<uint4*>(orow + col) = val;
                }
              }
            }
          }
        }
      }
    }
  }

#undef tO
#undef tP
#undef tS
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

// ------------------------------------------------------------------------------------------- head dim <= 64, 256 queries
// Same contract as attention_tc_kernel<1, 128>; one CTA = one (frame, head, 256-query tile).  The source-level profile of
// the 128-query kernel at the config-2 level-0 shape (profiles/r2_attention_q256.md) showed its softmax warps waiting for
// S = Q K^T 16 % of their time and one of the two groups idle during the merge-and-store tail: with split-KV inside the
// CTA the single MMA thread serves the groups strictly in turn (QK(j+2) for one group, PV(j), then the other group), so
// whichever group runs ahead waits for the other, and the merge is the work of group 0 alone.  Here the two softmax
// groups own DIFFERENT query tiles (rows 0-127 / 128-255) and walk the SAME key tiles:
//   * every K / V tile is staged once for 256 queries (half the TMA / L2 -> shared traffic per query);
//   * each group has its own MMA-issuing warp (warps 1 and 10): QK_g(j+1) goes out the moment group g has pulled S_g(j)
//     out of tensor memory and PV_g(j) the moment its P is stored -- the groups are coupled only through the depth of
//     the K / V rings (a stage is released when both issuers' MMAs on it have retired);
//   * no merge: each group normalises and stores its own 128 rows.
// Tensor memory: S_0, S_1 (2 x 128 columns), O_0, O_1 (2 x 64), P_0, P_1 (2 x 64 packed bf16 pairs) = 512 columns.
constexpr int NUM_THREADS_Q256 = 352;   // TMA warp, MMA warp of group 0, 2 x 4 softmax warps, MMA warp of group 1

// LMMA: the softmax row sums l = sum_k P[row, k] are produced by the tensor cores -- after P_g(j) V(j) the group's issuer adds
//       P_g(j) x ones (N = 16, B = a 2 KB shared-memory tile of bf16 1.0) into columns 48-63 of O_g, which P V at head dim
//       <= 48 (N = 48) leaves free -- instead of 16 FADD2 per 32 scores in the softmax warps; the lazy rescale of O_g covers
//       those columns, the epilogue reads l from column 48.  The normalisation then uses exactly the bf16 weights the
//       numerator was accumulated with.  Head dims 49-64 (P V at N = 64) keep the register sums.
// POLY: score pairs of every four that take the FMA-pipe exponential.  (Measured without gain and removed: S handed back to
//       the MMA warp a quarter into the tile, suspend-time hints on the waits of the TMA / MMA threads: r2_ab_flags.md.)
template <bool LMMA, int POLY>
__global__ void __launch_bounds__(NUM_THREADS_Q256, 1)
attention_tc_q256_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                         const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmK2,
                         const __grid_constant__ CUtensorMap tmV2, const AttnArgs args) {
  constexpr int BN = 128, DPAD = 64;
  constexpr int KS = 4, VS = 4;
  constexpr int Q_BYTES = BQ * DPAD * 2;         // one 128-query tile
  constexpr int KV_BYTES = BN * DPAD * 2;        // one K (or V) stage
  constexpr int TM_S = 0, TM_O = 2 * BN, TM_P = TM_O + 2 * DPAD;
  constexpr uint32_t IDESC_QK = idesc_bf16(BN, false);
  const int qk_steps = (args.d + 15) >> 4;       // real k-steps of Q K^T; P V runs at N = 16 * qk_steps columns
  const uint32_t IDESC_PV = idesc_bf16(qk_steps << 4, true);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem;                            // two query tiles
  uint8_t* sK = sQ + 2 * Q_BYTES;
  uint8_t* sV = sK + KS * KV_BYTES;
  uint8_t* sOnes = sV + VS * KV_BYTES;           // LMMA: 16 rows x 128 B of bf16 1.0 (any operand layout reads ones)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sOnes + 2048);
  uint64_t* q_full = bars;                       // 1
  uint64_t* k_full = bars + 1;                   // KS
  uint64_t* k_empty = k_full + KS;               // KS, two arrivals: one commit per issuer
  uint64_t* v_full = k_empty + KS;               // VS
  uint64_t* v_empty = v_full + VS;               // VS, two arrivals
  uint64_t* s_full = v_empty + VS;               // 2 (per group)
  uint64_t* s_empty = s_full + 2;                // 2
  uint64_t* p_full = s_empty + 2;                // 2
  uint64_t* p_empty = p_full + 2;                // 2
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(p_empty + 2);

  const int warp = uniform_warp_index(), lane = threadIdx.x & 31;
  // frames with the second key segment (twice the key tiles) first: see attention_tc_kernel
  const int n = (args.has_seg2 && args.seg2_index) ? (int)(gridDim.z - 1 - blockIdx.z) : (int)blockIdx.z;
  const int h = blockIdx.y, q0 = blockIdx.x * (2 * BQ);

  if (threadIdx.x == 0) {
    mbar_init(q_full, 1);
    for (int s = 0; s < KS; ++s) { mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 2); }
    for (int s = 0; s < VS; ++s) { mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 2); }
    for (int g = 0; g < 2; ++g) {
      mbar_init(&s_full[g], 1); mbar_init(&s_empty[g], 4); mbar_init(&p_full[g], 4); mbar_init(&p_empty[g], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  if constexpr (LMMA) {
    for (int i = threadIdx.x; i < 512; i += NUM_THREADS_Q256) reinterpret_cast<uint32_t*>(sOnes)[i] = 0x3F803F80u;
    fence_async_smem();                           // generic-proxy stores -> visible to the tensor-core (async) proxy
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_prologue();

  int seg2 = -1;
  if (args.has_seg2) seg2 = args.seg2_index ? args.seg2_index[n] : 0;
  const int tiles1 = (args.Lk + BN - 1) / BN;
  const int tiles2 = seg2 >= 0 ? (args.Lk2 + BN - 1) / BN : 0;
  const int num_tiles = tiles1 + tiles2;
  if (warp == 0) {
    if (elect_one_sync()) {
      // ===================== TMA producer =====================
      mbar_expect_tx(q_full, 2 * Q_BYTES);
      tma_load_3d(sQ, &tmQ, q_full, 0, h, n * args.Lq + q0);
      tma_load_3d(sQ + Q_BYTES, &tmQ, q_full, 0, h, n * args.Lq + q0 + BQ);   // past the tensor: zero fill
      auto tile_row = [&](int j) { return j >= tiles1 ? seg2 * args.Lk2 + (j - tiles1) * BN : n * args.Lk + j * BN; };
      auto load_k = [&](int j) {
        const int st = j % KS;
        mbar_wait(&k_empty[st], ((j / KS) & 1) ^ 1);
        mbar_expect_tx(&k_full[st], KV_BYTES);
        tma_load_3d(sK + st * KV_BYTES, j >= tiles1 ? &tmK2 : &tmK, &k_full[st], 0, h, tile_row(j));
      };
      auto load_v = [&](int j) {
        const int st = j % VS;
        mbar_wait(&v_empty[st], ((j / VS) & 1) ^ 1);
        mbar_expect_tx(&v_full[st], KV_BYTES);
        tma_load_3d(sV + st * KV_BYTES, j >= tiles1 ? &tmV2 : &tmV, &v_full[st], 0, h, tile_row(j));
      };
      // Each issuer retires QK(j+1) before PV(j): K(i + KS) can be requested once both QK(i) are done, V(i + VS) once
      // both PV(i) are -- in that order neither ring waits behind the other.
      for (int j = 0; j < KS && j < num_tiles; ++j) load_k(j);
      for (int j = 0; j < VS && j < num_tiles; ++j) load_v(j);
      for (int i = 0; i < num_tiles; ++i) {
        if (i + KS < num_tiles) load_k(i + KS);
        if (i + VS < num_tiles) load_v(i + VS);
      }
    }
  } else if (warp == 1 || warp == 10) {
    if (elect_one_sync()) {
      // ===================== MMA issuer of group g =====================
      const int g = warp == 1 ? 0 : 1;
      const uint32_t tS = tmem_base + TM_S + g * BN, tOg = tmem_base + TM_O + g * DPAD, tPg = tmem_base + TM_P + g * (BN / 2);
      const uint64_t dQ = desc_kmajor(smem_u32(sQ + g * Q_BYTES)), dOnes = desc_kmajor(smem_u32(sOnes));
      constexpr uint32_t IDESC_L = idesc_bf16(16, false);
      auto issue_qk = [&](int j) {
        const int st = j % KS;
        mbar_wait(&k_full[st], (j / KS) & 1);
        mbar_wait(&s_empty[g], (j & 1) ^ 1);               // group g has S(j - 1) in registers
        tc_fence_after();
        const uint64_t dK = desc_kmajor(smem_u32(sK + st * KV_BYTES));
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (k < qk_steps) umma(tS, dQ + ((k * 32) >> 4), dK + ((k * 32) >> 4), IDESC_QK, k != 0);
        umma_commit(&k_empty[st]);
        umma_commit(&s_full[g]);
      };
      auto issue_pv = [&](int j) {
        const int st = j % VS;
        mbar_wait(&v_full[st], (j / VS) & 1);
        mbar_wait(&p_full[g], j & 1);
        tc_fence_after();
        const uint64_t dV = desc_mnmajor(smem_u32(sV + st * KV_BYTES), BN * 128);
#pragma unroll
        for (int k = 0; k < BN / 16; ++k) umma_ts(tOg, tPg + k * 8, dV + k * (2048 >> 4), IDESC_PV, (j | k) != 0);
        umma_commit(&v_empty[st]);
        if constexpr (LMMA) {                        // row sums: O_g[:, 48..63] += P_g(j) x ones
#pragma unroll
          for (int k = 0; k < BN / 16; ++k) umma_ts(tOg + 48, tPg + k * 8, dOnes, IDESC_L, (j | k) != 0);
        }
        umma_commit(&p_empty[g]);
      };
      mbar_wait(q_full, 0);
      issue_qk(0);
      for (int j = 0; j < num_tiles; ++j) {
        if (j + 1 < num_tiles) issue_qk(j + 1);
        issue_pv(j);
      }
    }
  } else {
    // ===================== softmax group g: query rows g * 128 .. g * 128 + 127 =====================
    const int g = (warp - 2) >> 2;
    const int lane_grp = warp & 3;
    const int row = lane_grp * 32 + lane;            // query row inside the group's tile == TMEM lane
    const uint32_t lane_addr = (uint32_t)(lane_grp * 32) << 16;
    const float c = args.scale_log2e;
    float m_ref = -INFINITY;   // reference maximum the stored O_g / l are relative to (raw score units)
    float l_run = 0.f;
    const uint32_t tS = tmem_base + TM_S + g * BN + lane_addr;
    const uint32_t tO = tmem_base + TM_O + g * DPAD + lane_addr;
    const uint32_t tP = tmem_base + TM_P + g * (BN / 2) + lane_addr;

    for (int j = 0; j < num_tiles; ++j) {
      const bool second = j >= tiles1;
      const int seg_len = second ? args.Lk2 : args.Lk;
      const int k0 = (second ? (j - tiles1) : j) * BN;
      const int valid = min(BN, seg_len - k0);       // keys of this tile that exist
      mbar_wait(&s_full[g], j & 1);
      tc_fence_after();
      // online softmax in 32-column chunks, the TMEM load of chunk i + 1 in flight during chunk i (see attention_tc_kernel)
      uint32_t sa[32], sb[32];
      uint32_t pk[BN / 2];
      float alpha_tile = 1.f;
      float lsum[4] = {0.f, 0.f, 0.f, 0.f};
      auto chunk = [&](uint32_t (&sc)[32], const int ci) {
        if (valid < BN) {   // partial last tile of a segment (warp-uniform branch)
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (ci * 32 + i >= valid) sc[i] = 0xff800000u;   // -inf
        }
        float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
#pragma unroll
          for (int e = 0; e < 4; ++e) mx4[e] = fmaxf(mx4[e], __uint_as_float(sc[i + e]));
        }
        const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
        if ((mx - m_ref) * c > RESCALE_THRESHOLD) {              // true for the first chunk ever (m_ref = -inf)
          const float a = ex2_approx((m_ref - mx) * c);          // 0 the first time
          m_ref = mx;
          alpha_tile *= a;
          if constexpr (!LMMA) {
            l_run *= a;
#pragma unroll
            for (int e = 0; e < 4; ++e) lsum[e] *= a;
          }
          const __nv_bfloat162 a2 = __float2bfloat162_rn(a);
#pragma unroll
          for (int i = 0; i < BN / 2; ++i) {
            if (i < ci * 16) {                                    // chunks of this tile that are already packed
              __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&pk[i]);
              v = __hmul2(v, a2);
              pk[i] = *reinterpret_cast<uint32_t*>(&v);
            }
          }
        }
        const float mc = m_ref * c;
#pragma unroll
        for (int i = 0; i < 32; i += 8) exp8_pack_sum<POLY, !LMMA>(sc + i, c, mc, pk + ci * 16 + i / 2, lsum);
      };
      tmem_ld32(tS, sa);
      tmem_ld_wait();
      pin32(sa);
#pragma unroll
      for (int ci = 0; ci < BN / 32; ci += 2) {
        tmem_ld32(tS + (ci + 1) * 32, sb);                        // in flight during chunk ci
        chunk(sa, ci);
        tmem_ld_wait();
        pin32(sb);
        if (ci + 2 < BN / 32) {
          tmem_ld32(tS + (ci + 2) * 32, sa);                      // in flight during chunk ci + 1
        } else {
          tc_fence_before();                                      // S is in registers: hand the buffer back
          __syncwarp();
          if (lane == 0) mbar_arrive(&s_empty[g]);
        }
        chunk(sb, ci + 1);
        if (ci + 2 < BN / 32) {
          tmem_ld_wait();
          pin32(sa);
        }
      }
      if constexpr (!LMMA) l_run += (lsum[0] + lsum[1]) + (lsum[2] + lsum[3]);

      if (j > 0) {
        mbar_wait(&p_empty[g], (j - 1) & 1);   // this group's previous PV retired: P buffer reusable, O_g stable
        tc_fence_after();
        if (__any_sync(0xffffffffu, alpha_tile != 1.f)) {
#pragma unroll
          for (int cc = 0; cc < DPAD / 32; ++cc) {
            uint32_t o[32];
            tmem_ld32(tO + cc * 32, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha_tile);
            tmem_st32(tO + cc * 32, o);
          }
          tmem_st_wait();
        }
      }
#pragma unroll
      for (int i = 0; i < BN / 64; ++i) tmem_st32(tP + i * 32, pk + i * 32);   // lane = query row, column = key pair
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[g]);
    }

    // ---- normalise and store this group's rows (only the first d columns are real)
    if (num_tiles > 0) {
      mbar_wait(&p_empty[g], (num_tiles - 1) & 1);
      tc_fence_after();
      const int q = q0 + g * BQ + row;
      bf16* orow = args.out + ((int64_t)n * args.Lq + q) * args.ldo + h * args.d;
      uint32_t o[2][32];
      tmem_ld32(tO, o[0]);
      if (LMMA || args.d > 32) tmem_ld32(tO + 32, o[1]);     // warp-uniform
      tmem_ld_wait();
      const float inv = 1.f / (LMMA ? __uint_as_float(o[1][16]) : l_run);   // column 48 = sum_k P[row, k]
      if (q < args.Lq) {
#pragma unroll
        for (int cc = 0; cc < DPAD / 32; ++cc) {
#pragma unroll
          for (int gg = 0; gg < 4; ++gg) {
            const int col = cc * 32 + gg * 8;
            if (col < args.d) {
              uint4 val;
              val.x = pack_bf16(__uint_as_float(o[cc][gg * 8 + 0]) * inv, __uint_as_float(o[cc][gg * 8 + 1]) * inv);
              val.y = pack_bf16(__uint_as_float(o[cc][gg * 8 + 2]) * inv, __uint_as_float(o[cc][gg * 8 + 3]) * inv);
              val.z = pack_bf16(__uint_as_float(o[cc][gg * 8 + 4]) * inv, __uint_as_float(o[cc][gg * 8 + 5]) * inv);
              val.w = pack_bf16(__uint_as_float(o[cc][gg * 8 + 6]) * inv, __uint_as_float(o[cc][gg * 8 + 7]) * inv);
              *reinterpret_cast<uint4*>(orow + col) = val;
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

// ------------------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int encode_qkv_map(mmgt_ctx* ctx, CUtensorMap* map, const void* base, int d, int heads, int64_t rows, int64_t ld, int box_rows) {
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
  EncodeTiledFn fn = reinterpret_cast<EncodeTiledFn>(ctx->encode_tiled);
  cuuint64_t dims[3] = {(cuuint64_t)d, (cuuint64_t)heads, (cuuint64_t)rows};
  cuuint64_t strides[2] = {(cuuint64_t)d * 2, (cuuint64_t)ld * 2};
  cuuint32_t box[3] = {64, 1, (cuuint32_t)box_rows};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    mmgt_set_error("attention: cuTensorMapEncodeTiled failed (%d) d=%d heads=%d rows=%lld ld=%lld", (int)r, d, heads,
                   (long long)rows, (long long)ld);
    return MMGT_E_INVALID;
  }
  return 0;
}

template <int DCH, int BN, bool PACKED>
int launch_attn(mmgt_ctx* ctx, const CUtensorMap* maps, const AttnArgs& a, cudaStream_t st) {
  constexpr bool p_tmem = (2 * BN + 2 * 64 * DCH + BN) <= 512;
  constexpr int smem = BQ * 64 * DCH * 2 + (k_stages(DCH) + v_stages(DCH)) * BN * 64 * DCH * 2 + (p_tmem ? 0 : 2 * BQ * BN * 2) + BQ * 8 + 1024 + 256;
  static_assert(smem <= 232448, "shared memory budget");
  MMGT_CUDA_OK(mmgt_smem_optin(ctx, attention_tc_kernel<DCH, BN, PACKED>, smem));
  dim3 grid((a.Lq + BQ - 1) / BQ, a.heads, a.N);
  MMGT_CUDA_OK(mmgt_launch(ctx, attention_tc_kernel<DCH, BN, PACKED>, grid, dim3(NUM_THREADS), smem, st, maps[0], maps[1], maps[2], maps[3],
                           maps[4], a));
  MMGT_LAUNCH_OK(ctx);
  return 0;
}

template <bool LMMA, int POLY>
int launch_attn_q256(mmgt_ctx* ctx, const CUtensorMap* maps, const AttnArgs& a, cudaStream_t st) {
  constexpr int smem = 2 * BQ * 64 * 2 + (4 + 4) * 128 * 64 * 2 + 2048 + 1024 + 256;
  MMGT_CUDA_OK(mmgt_smem_optin(ctx, attention_tc_q256_kernel<LMMA, POLY>, smem));
  dim3 grid((a.Lq + 2 * BQ - 1) / (2 * BQ), a.heads, a.N);
  MMGT_CUDA_OK(mmgt_launch(ctx, attention_tc_q256_kernel<LMMA, POLY>, grid, dim3(NUM_THREADS_Q256), smem, st, maps[0], maps[1],
                           maps[2], maps[3], maps[4], a));
  MMGT_LAUNCH_OK(ctx);
  return 0;
}

int launch_attn_persist(mmgt_ctx* ctx, const CUtensorMap* maps, const AttnArgs& a, cudaStream_t st) {
  constexpr int smem = 2 * BQ * 64 * 2 + (4 + 3) * 128 * 64 * 2 + 2 * BQ * 8 + 1024 + 256;
  MMGT_CUDA_OK(mmgt_smem_optin(ctx, attention_tc_persist_kernel<128>, smem));
  const int q_tiles = (a.Lq + BQ - 1) / BQ;
  const int64_t items = (int64_t)q_tiles * a.heads * a.N;
  const int grid = (int)std::min<int64_t>(items, ctx->attn_persist >= 2 ? ctx->attn_persist : ctx->num_sms);
  MMGT_CUDA_OK(mmgt_launch(ctx, attention_tc_persist_kernel<128>, dim3(grid), dim3(NUM_THREADS), smem, st, maps[0], maps[1], maps[2],
                           maps[3], maps[4], a, q_tiles, (int)items));
  MMGT_LAUNCH_OK(ctx);
  return 0;
}

int launch_attn64(mmgt_ctx* ctx, const CUtensorMap* maps, const AttnArgs& a, cudaStream_t st) {
  constexpr int smem = BQ * 64 * 2 + (4 + 3) * 128 * 64 * 2 + BQ * 8 + 1024 + 256;
  MMGT_CUDA_OK(mmgt_smem_optin(ctx, attention_tc64_kernel<128>, smem));
  dim3 grid((a.Lq + BQ - 1) / BQ, a.heads, a.N);
  MMGT_CUDA_OK(mmgt_launch(ctx, attention_tc64_kernel<128>, grid, dim3(NUM_THREADS), smem, st, maps[0], maps[1], maps[2], maps[3],
                           maps[4], a));
  MMGT_LAUNCH_OK(ctx);
  return 0;
}

}  // namespace

bool mmgt_attention_tc_supported(const mmgt_ctx* ctx, const mmgt_attention_params* p) {
  (void)ctx;
  if (p->dtype != MMGT_BF16) return false;
  if (p->d % 8 || p->d > 192 || p->d < 8) return false;
  if (p->Lq < 64 || p->N > 65535 || p->heads > 65535) return false;
  if (p->ldq % 8 || p->ldk % 8 || p->ldv % 8 || p->ldo % 8) return false;
  if (p->kv_batch_stride && (p->kv_batch_stride != (int64_t)p->Lk * p->ldk || p->ldk != p->ldv)) return false;
  if (!aligned16(p->q) || !aligned16(p->k) || !aligned16(p->v) || !aligned16(p->out)) return false;
  if (p->k2 && (p->ldk2 % 8 || p->ldv2 % 8 || !aligned16(p->k2) || !aligned16(p->v2) || p->B2 < 1)) return false;
  return true;
}

int mmgt_attention_tc(mmgt_ctx* ctx, const mmgt_attention_params* p, cudaStream_t st) {
  const int dch = (p->d + 63) / 64;
  const int bn = dch == 3 ? 64 : 128;
  CUtensorMap maps[5];
  int rc;
  if ((rc = encode_qkv_map(ctx, &maps[0], p->q, p->d, p->heads, (int64_t)p->N * p->Lq, p->ldq, BQ))) return rc;
  if ((rc = encode_qkv_map(ctx, &maps[1], p->k, p->d, p->heads, (int64_t)p->N * p->Lk, p->ldk, bn))) return rc;
  if ((rc = encode_qkv_map(ctx, &maps[2], p->v, p->d, p->heads, (int64_t)p->N * p->Lk, p->ldv, bn))) return rc;
  const void* k2 = p->k2 ? p->k2 : p->k;
  const void* v2 = p->k2 ? p->v2 : p->v;
  const int64_t rows2 = p->k2 ? (int64_t)p->B2 * p->Lk2 : (int64_t)p->N * p->Lk;
  if ((rc = encode_qkv_map(ctx, &maps[3], k2, p->d, p->heads, rows2, p->k2 ? p->ldk2 : p->ldk, bn))) return rc;
  if ((rc = encode_qkv_map(ctx, &maps[4], v2, p->d, p->heads, rows2, p->k2 ? p->ldv2 : p->ldv, bn))) return rc;
  AttnArgs a{};
  a.N = p->N; a.Lq = p->Lq; a.Lk = p->Lk; a.Lk2 = p->k2 ? p->Lk2 : 0; a.heads = p->heads; a.d = p->d;
  a.seg2_index = p->seg2_index; a.has_seg2 = p->k2 != nullptr;
  a.out = (bf16*)p->out; a.ldo = p->ldo;
  a.scale_log2e = p->scale * 1.4426950408889634f;
  const bool pk = ctx->attn_packed != 0;   // flag 16: FFMA2 / FADD2 in the softmax loops
  if (dch == 1) {
    if (ctx->attn_v2) return launch_attn64(ctx, maps, a, st);
    const int64_t items = (int64_t)((p->Lq + BQ - 1) / BQ) * p->heads * p->N;
    if (items < (1ll << 31) && (ctx->attn_persist >= 2 || (ctx->attn_persist == 1 && items > ctx->num_sms)))
      return launch_attn_persist(ctx, maps, a, st);
    // 256 queries per CTA once both softmax groups have rows of their own (flag 15)
    if (ctx->attn_q256 && p->Lq > BQ) {
      const bool lmma = p->d <= 48;        // P V at N <= 48 leaves columns 48-63 of O for the row sums
      switch (ctx->attn_q256) {            // A/B variants: share of FMA-pipe exponentials, row sums by the tensor cores
        case 1: return launch_attn_q256<false, 0>(ctx, maps, a, st);
        case 4: return launch_attn_q256<false, 2>(ctx, maps, a, st);
        case 7: return lmma ? launch_attn_q256<true, 1>(ctx, maps, a, st) : launch_attn_q256<false, 1>(ctx, maps, a, st);
        case 8: return lmma ? launch_attn_q256<true, 0>(ctx, maps, a, st) : launch_attn_q256<false, 0>(ctx, maps, a, st);
        case 9: return lmma ? launch_attn_q256<true, 2>(ctx, maps, a, st) : launch_attn_q256<false, 2>(ctx, maps, a, st);
        default: return launch_attn_q256<false, 1>(ctx, maps, a, st);
      }
    }
    return pk ? launch_attn<1, 128, true>(ctx, maps, a, st) : launch_attn<1, 128, false>(ctx, maps, a, st);
  }
  if (dch == 2) return pk ? launch_attn<2, 128, true>(ctx, maps, a, st) : launch_attn<2, 128, false>(ctx, maps, a, st);
  return pk ? launch_attn<3, 64, true>(ctx, maps, a, st) : launch_attn<3, 64, false>(ctx, maps, a, st);
}
