/*
 * mmgt_b200.h -- C ABI of the B200-native kernels behind MMGT's stage-2 denoising hot path.
 *
 * Drop-in boundary: the reference is 100 % Python/PyTorch and has no FFI of
 * its own; the operators below are exactly the implicit library kernels its hot path dispatches
 * (SURVEY.md section 2.3), one entry point per fused operator.  Each declaration cites the reference
 * call site it replaces (paths relative to the reference checkout).  The Python host in
 * mmgt_b200/ binds them with ctypes (see INTEGRATION.md for the stub a maintainer would add).
 *
 * Conventions
 *   - every entry point returns int: 0 = OK, <0 = invalid argument (MMGT_E_*), >0 = cudaError_t.
 *   - nothing here allocates device memory, synchronises the stream, or throws.
 *   - all pointers are DEVICE pointers unless the name ends in _host.
 *   - activations are channels-last: a frame batch is (N, T, C) row-major with N = batch*frames,
 *     T = H*W tokens, C channels; "rows" = N*T.
 *   - dtype: MMGT_F32 or MMGT_BF16 selects the storage type of activations / GEMM weights.
 *     bias / norm affine / scale vectors are always float32.  Accumulation is always float32.
 *   - stream is a cudaStream_t passed as void*.
 */
#ifndef MMGT_B200_H
#define MMGT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MMGT_F32 0
#define MMGT_BF16 1

#define MMGT_E_INVALID (-1)      /* bad shape / null pointer / unsupported combination */
#define MMGT_E_ALIGN (-2)        /* pointer or leading dimension not aligned as required */
#define MMGT_E_UNSUPPORTED (-3)  /* valid request but no kernel for it (never silently emulated) */
#define MMGT_E_NODRIVER (-4)     /* cuTensorMapEncodeTiled could not be resolved */

#if defined(__GNUC__)
#define MMGT_API __attribute__((visibility("default")))
#else
#define MMGT_API
#endif

typedef struct mmgt_ctx mmgt_ctx;

/* Library / context ------------------------------------------------------------------------- */
MMGT_API int mmgt_abi_version(void);
/* Creates the per-device context (SM count, opt-in shared memory, driver entry points). */
MMGT_API int mmgt_ctx_create(mmgt_ctx** out, int device);
MMGT_API int mmgt_ctx_destroy(mmgt_ctx* ctx);
/* Last error message of the calling thread (static storage, never NULL). */
MMGT_API const char* mmgt_last_error(void);
/* flag 0: enable (1) / disable (0) the tcgen05 tensor-core kernels for bf16 (default 1).
 * flag 1: number of kernels launched through this context since creation (read with value<0).
 * flag 2: enable (1) / disable (0) the weight-stationary tensor-core GEMM variant for small K (default 1).
 * flag 3: enable (1) / disable (0) programmatic dependent launch: every kernel is launched with programmatic stream
 *         serialization and waits (griddepcontrol.wait) before its first global access, so its scheduling and
 *         prologue overlap the tail of the previous kernel; stream semantics are unchanged.
 * flag 4: strict tensor-core mode (default 0): a bf16 GEMM / convolution / attention request that no tcgen05 kernel
 *         covers returns MMGT_E_UNSUPPORTED instead of running on the CUDA-core kernels (shape cliffs become errors).
 * flag 5: number of bf16 operator calls that ran on a CUDA-core kernel while flag 0 was on (read with value<0).
 * flag 6: 1 = the tensor-core GEGLU epilogue evaluates GELU through erf (Abramowitz-Stegun 7.1.26) instead of the
 *         default logistic-polynomial fit (<= 8.2e-4 relative, 0.4 bf16 half-ulps); A/B switch for tests.
 * flag 7: GroupNorm as two kernels, statistics then normalise (default 1); 0 = the single kernel with a per-frame
 *         arrival barrier (grid limited to co-resident CTAs).  A/B switch.
 * flag 8: stride-2 and upsampling 3x3 convolutions as implicit GEMMs (TMA traversal stride 2; four sub-pixel 2x2-tap
 *         convolutions) (default 1); 0 = stage an im2col matrix and run the plain GEMM.  A/B switch.
 * flag 9: head dim <= 64 attention on the kernel with three rotating S buffers and P aliased over S; default 0 = the
 *         two-buffer kernel, which measured 1.5 % faster per DDIM step (profiles/r2_ab_flags.md).  A/B switch.
 * flag 10: temporal attention with head dim <= 80 on the row-coalesced kernel (one CTA per (batch, pixel), whole q|k|v rows
 *         through double-buffered cp.async) (default 1); 0 = one warp per (batch, pixel, head).  A/B switch.
 * flag 11: tensor-core GEMM / conv launches whose epilogue needs only bias, a per-tile row bias, GEGLU or a residual take
 *         a kernel instance with that epilogue compiled straight-line (default 1); 0 = always the general epilogue
 *         with its run-time option branches.  A/B switch.
 * flag 12: the specialised residual epilogues write their output tiles through TMA stores from swizzled shared-memory
 *         boxes (default 1) where a tile is 128 consecutive output rows; 0 = one 32-byte store per lane and 16-column
 *         chunk (32 different lines per warp instruction), as the non-residual epilogues always do.  A/B switch.
 * flag 13: residual epilogues of the streaming tensor-core GEMM / conv kernels feed the residual through the tensor
 *         cores -- extra k-blocks [residual tile | identity], exact in fp32 -- instead of loading a row per lane
 *         (default 1); 0 = per-lane loads (+ flag 12).  The context owns the 256 x 256 bf16 identity (128 KB, allocated
 *         in mmgt_ctx_create, freed in mmgt_ctx_destroy).  A/B switch.
 * flag 14: head dim <= 64 attention on persistent CTAs (default 0 -- measured slower; 1 = when items > SMs): one CTA per SM walks the (frame, head, query tile)
 *         items, the next item's loads and first Q K^T overlapping the merge-and-store tail of the current one; 0 = one
 *         item per CTA; n >= 2 = always, on at most n CTAs (tests).  Same arithmetic, bit-identical results.  A/B switch.
 * flag 15: head dim <= 64 attention with 256 queries per CTA (default 5): the two softmax groups own different query
 *         tiles and walk the same key tiles, each with its own MMA-issuing warp; no split-KV merge.  0 = 128-query CTAs
 *         whose groups take alternate key tiles.  Variants of the 256-query kernel: 1 = every exponential on MUFU.EX2;
 *         5 = one score pair in four through a degree-3 polynomial on the FMA pipe (7.5e-5 relative, far inside the bf16
 *         rounding of P); 4 = two pairs in four (measured slower); 8 / 7 / 9 = 1 / 5 / 4 with the softmax row sums taken
 *         from the tensor cores (P x ones into accumulator columns 48-63, head dim <= 48) instead of register sums.
 *         A/B switch.
 * flag 17: LayerNorm grid: 0 = up to 16 blocks per SM (default), 1 = one exact wave of persistent blocks, 2 = the same with
 *         the next row of a lane group prefetched.  Bit-identical results.  A/B switch.
 * flag 16: attention softmax loops on packed fp32 pairs (fma.rn.f32x2 / add.rn.f32x2; default 1).  Same IEEE operations
 *         as the scalar form: bit-identical results.  A/B switch. */
MMGT_API int64_t mmgt_ctx_flag(mmgt_ctx* ctx, int flag, int64_t value);

/* Layout ------------------------------------------------------------------------------------- */
/* (B,C,F,H,W) float32|bf16 -> (B*F, H*W, C) in `dtype`, optionally adding a second NCFHW tensor.
 * Replaces the rearranges at resnet.py:13, transformer_3d.py:158 and the pose add unet_3d.py:517-519. */
MMGT_API int mmgt_ncfhw_to_tokens(mmgt_ctx*, const void* src, const void* add_or_null, void* dst, int B, int C, int F,
                         int H, int W, int src_dtype, int dst_dtype, void* stream);
/* (B*F, H*W, src_ld >= C) -> (B,C,F,H,W); inverse of the above (resnet.py:15, transformer_3d.py:264).
 * src_ld is the channel stride of the token tensor (0 = C); conv_out is computed with padded output channels. */
MMGT_API int mmgt_tokens_to_ncfhw(mmgt_ctx*, const void* src, void* dst, int B, int C, int F, int H, int W, int src_ld,
                         int src_dtype, int dst_dtype, void* stream);

/* Normalisation ------------------------------------------------------------------------------ */
/* GroupNorm over (T, C/groups) per frame on a channels-last tensor, optional fused SiLU.  The input may
 * be the virtual channel-concat [x1 (C1) || x2 (C2)] (x2 may be NULL, C2 = 0); the output is the
 * concatenated normalised tensor (N,T,C1+C2).  stats_ws: >= N*groups*2 + ceil(N/2) doubles of scratch
 * (per-group sums, then one 32-bit arrival counter per frame; zeroed by the call itself).
 * Replaces InflatedGroupNorm/nn.GroupNorm + F.silu (resnet.py:20-28,220-221,231-237;
 * transformer_3d.py:174; motion_module.py:156; unet_3d.py:618-619) and torch.cat
 * (unet_3d_blocks.py:894,1057). */
MMGT_API int mmgt_groupnorm(mmgt_ctx*, const void* x1, const void* x2_or_null, void* y, const float* gamma,
                   const float* beta, double* stats_ws, int N, int T, int C1, int C2, int groups, float eps,
                   int silu, int dtype, void* stream);
/* LayerNorm over C per row (eps as given); optional positional-encoding add: y += pe[(row/T) % F, :]
 * (pe is (max_len, C) float32, NULL = none).  Replaces nn.LayerNorm (attention.py:331-362,576-644;
 * motion_module.py:228-244) and PositionalEncoding.forward (motion_module.py:275-277,365-366). */
MMGT_API int mmgt_layernorm(mmgt_ctx*, const void* x, void* y, const float* gamma, const float* beta,
                   const float* pe_or_null, int64_t rows, int C, int T, int F, float eps, int dtype,
                   void* stream);

/* (mean, rstd) of every row of x (rows, C) with leading dimension ld (0 = C): stats (rows, 2) float32.  The LayerNorm
 * itself is then applied by the GEMM that consumes x (mmgt_gemm_params.rowstats / colsum), so the normalised tensor
 * is never written (attention.py:331-362,576-644; motion_module.py:228-244). */
MMGT_API int mmgt_row_stats(mmgt_ctx*, const void* x, float* stats, int64_t rows, int C, int64_t ld, float eps, int dtype,
                            void* stream);

/* Frame-shard <-> token-shard row exchange (multi-GPU, SURVEY.md section 8e level 3) ------------------
 * k ranks share one context window: each holds F/k of the F frames ("frame-sharded", rows (b, f_loc, t)).
 * Around every motion module the rows switch to "token-sharded" (rows (b, f, t_loc): all F frames of T/k pixels)
 * -- the (b f) d c <-> (b d) f c rearranges of motion_module.py:361-363,386 turned into an all-to-all.  The exchange
 * is written by the PRODUCING kernel: destination rows are stored straight into the peers' receive buffers over
 * NVLink (peer_base[s] = receive buffer of shard s mapped into this process; peer_base[my] is the local one).
 *   direction 1 (frame -> token): source row m = (b*F/k + f_loc)*T + t      -> shard t / (T/k),
 *                                 destination row (b*F + my*F/k + f_loc)*(T/k) + t % (T/k)
 *   direction 2 (token -> frame): source row m = (b*F + f)*(T/k) + t_loc    -> shard f / (F/k),
 *                                 destination row (b*F/k + f % (F/k))*T + my*(T/k) + t_loc */
#define MMGT_MAX_PEERS 8
typedef struct {
  void* peer_base[MMGT_MAX_PEERS];
  int k;          /* shards (2..MMGT_MAX_PEERS) */
  int my;         /* this rank's shard */
  int direction;  /* 1 or 2 */
  int B, F, T;    /* samples, frames per sample of the WHOLE window, tokens per frame; F % k == 0, T % k == 0 */
  int64_t ld;     /* leading dimension (elements) of the destination rows */
} mmgt_row_exchange;

/* GEMM / convolution --------------------------------------------------------------------------- */
typedef struct {
  const void* A;        /* (M,K) row-major, leading dim lda (elements) */
  const void* W;        /* (N,K) row-major (nn.Linear / 1x1-conv weight), leading dim ldw */
  void* D;              /* (M,N) or (M,N/2) for GEGLU, leading dim ldd */
  const float* bias;    /* (N) or NULL */
  const float* rowscale;/* (M) or NULL: per-row multiplier applied after bias (MM-HAA mask gate) */
  const float* rowbias; /* (ceil(M/rows_per_group), N) or NULL: broadcast add (time embedding, CLIP) */
  const void* residual; /* (M,N_out) or NULL, leading dim ldr; may alias D */
  int64_t lda, ldw, ldd, ldr;
  int M, N, K;
  int rows_per_group;
  float alpha;          /* D = alpha*rowscale*(A W^T + bias) + rowbias + residual */
  int geglu_block;      /* 0 = off; else W rows are interleaved [value(gb) | gate(gb)]* and
                           D[m, j] = value * gelu_erf(gate), N_out = N/2 (diffusers GEGLU) */
  int dtype;            /* storage type of A, W, D, residual */
  int out_f32;          /* 1: D is float32 regardless of dtype (small-M vectors) */
  const mmgt_row_exchange* exchange; /* HOST pointer or NULL.  Non-NULL: the epilogue stores row m of the result to the
                           peer / row given by the exchange instead of D + m*ldd (D is ignored, may be NULL);
                           tensor-core path only (bf16), otherwise MMGT_E_UNSUPPORTED */
  int64_t ld_rowbias;   /* elements between rowbias rows (0 = N_out): lets one (B, sum of widths) time-embedding
                           projection serve every resnet as a column slice (resnet.py:226) */
  int rowbias_mod;      /* > 0: rowbias row = (m / rows_per_group) % rowbias_mod (the motion module's positional
                           table added per frame of every sample, motion_module.py:365-366) */
  const float* rowstats;/* (M,2) float32 [mean, rstd] or NULL.  Non-NULL fuses the LayerNorm that feeds this GEMM:
                           with W' = W diag(gamma), colsum[n] = sum_k W'[n,k] and bias' = bias + W beta the result is
                           rstd*(A W'^T - mean*colsum) + bias' = LN(A) W^T + bias (attention.py:331-362) */
  const float* colsum;  /* (N) float32, required with rowstats */
  int act;              /* 0 none, 1 SiLU, 2 ReLU: applied after bias / rowscale / rowbias, before the residual
                           (pose_guider.py:47-57, audio_proj.py:96) */
} mmgt_gemm_params;
/* Replaces nn.Linear / 1x1 nn.Conv2d + bias + residual adds + GEGLU (diffusers Attention.to_q/k/v/out,
 * FeedForward; transformer_3d.py:176,253; resnet.py:226,243; attention.py:730-767;
 * motion_module.py:161,172). */
MMGT_API int mmgt_gemm(mmgt_ctx*, const mmgt_gemm_params*, void* stream);
/* Widest N tile of the tensor-core kernel that divides N (0 = none: such weights run on the CUDA-core kernel).
 * A GEGLU weight must be row-interleaved with geglu_block = 16 to take the tensor-core path. */
MMGT_API int mmgt_gemm_tc_block_n(int N);

typedef struct {
  const void* x;        /* (N, H, W, Cin) channels-last */
  const void* w;        /* (Cout, 3, 3, Cin) "KRSC" repack of the (Cout,Cin,3,3) weight */
  void* y;              /* (N, Ho, Wo, Cout) */
  const float* bias;    /* (Cout) or NULL */
  const float* rowbias; /* (N/frames_per_group, Cout) or NULL: time-embedding add (resnet.py:226-229) */
  const void* residual; /* (N, Ho, Wo, Cout) or NULL; may alias y */
  int N, H, W, Cin, Cout;
  int stride;           /* 1 or 2 (Downsample3D, resnet.py:106-108), padding is always 1 */
  int upsample2x;       /* 1: x is nearest-upsampled x2 on the fly before the conv (Upsample3D, resnet.py:70-88) */
  int frames_per_group; /* frames sharing one rowbias row (= F) */
  int dtype;
  int64_t ld_rowbias;   /* elements between rowbias rows (0 = Cout) */
  int act;              /* 0 none, 1 SiLU, 2 ReLU (after bias / rowbias, before the residual) */
  const void* w_subpixel; /* upsample2x only, or NULL: (Cout, 4, 2, 2, Cin) pre-summed taps of the four output parities
                           (a, b) = (row, column) parity, index 2a + b; tap (ty, tx) of parity (a, b) reads input pixel
                           (h + ty - (a ? 0 : 1), w + tx - (b ? 0 : 1)).  With it the bf16 tensor-core path runs the
                           upsampling convolution as four 2x2-tap implicit GEMMs on the low-resolution input (2.25x fewer
                           FLOPs, nothing staged); without it the operator stages an im2col matrix. */
} mmgt_conv3x3_params;
/* Replaces InflatedConv3d k=3 (resnet.py:9-17) incl. the fused epilogue of ResnetBlock3D.
 * workspace: scratch of at least mmgt_conv3x3_workspace_bytes() bytes (may be NULL when that is 0;
 * only the stride-2 / upsampling tensor-core paths stage an im2col matrix). */
MMGT_API int64_t mmgt_conv3x3_workspace_bytes(mmgt_ctx*, const mmgt_conv3x3_params*);
MMGT_API int mmgt_conv3x3(mmgt_ctx*, const mmgt_conv3x3_params*, void* workspace, int64_t workspace_bytes, void* stream);

/* Attention ------------------------------------------------------------------------------------ */
typedef struct {
  const void* q;  /* (N, Lq, heads*d), row stride ldq */
  const void* k;  /* (N, Lk, heads*d), row stride ldk: per-frame keys   */
  const void* v;  /* (N, Lk, heads*d), row stride ldv                  */
  const void* k2; /* (B2, Lk2, heads*d) row stride ldk2: shared second key segment (reference bank) or NULL */
  const void* v2;
  const int32_t* seg2_index; /* (N) device: row of k2/v2 used by frame n, or -1 = first segment only
                                (the CFG uncond half, mutual_self_attention.py:168-188).  NULL with k2 => 0 */
  void* out;      /* (N, Lq, heads*d), row stride ldo */
  int64_t ldq, ldk, ldv, ldk2, ldv2, ldo;
  int64_t kv_batch_stride; /* elements between consecutive frames of k/v (0 = Lk*ldk) */
  int N, Lq, Lk, Lk2, heads, d;
  int B2;         /* rows (batch) of k2 / v2 */
  float scale;    /* d^-0.5 */
  int dtype;
} mmgt_attention_params;
/* softmax(q k^T * scale) v over [k ; k2].  Replaces diffusers AttnProcessor2_0 /
 * F.scaled_dot_product_attention for spatial self(+reference) attention and the audio cross-attention
 * (mutual_self_attention.py:157-188; attention.py:694,720-750). */
MMGT_API int mmgt_attention(mmgt_ctx*, const mmgt_attention_params*, void* stream);

/* MM-HAA audio cross-attention, three regions in one launch with the mask gate fused (attention.py:719-767).
 * For region r in {full, face, lip}, frame n, row t, head h:
 *   out[n*T+t, r*C + h*d + :] = gate_r * softmax_k(q3[n*T+t, r*C + h*d + :] . K_r[n, k, h] * softmax_scale) V_r[n, :, h]
 *   out[n*T+t, 3C + r] = gate_r,  out[n*T+t, 3C+3 .. 3C+7] = 0,   gate_r = mask[r][n*T+t] * scale[r]
 * with K_r = kv6[:, 2r*C : (2r+1)*C], V_r = kv6[:, (2r+1)*C : (2r+2)*C] (the fused to_k / to_v projections of the M
 * audio tokens of frame n).  One GEMM of `out` (K = 3C + 8) against [Wz_0 Wo_0 | Wz_1 Wo_1 | Wz_2 Wo_2 | Wz_r bo_r | 0]
 * then yields sum_r scale_r * zero_conv_r(mask_r * to_out_r(attn_r)) (+ bias, + residual) -- replacing, per layer,
 * three SDPA calls, three mask multiplies, six projections and the 3-way add (attention.py:719-767).  bf16 only. */
typedef struct {
  const void* q3;        /* (N*T, 3C), row stride ldq */
  const void* kv6;       /* (N*M, 6C), row stride ldkv */
  const float* mask[3];  /* (N*T) float32 each: full, face, lip motion masks at this level */
  float scale[3];        /* motion_scale per region (1 when it does not reach MM-HAA) */
  void* out;             /* (N*T, ldo >= 3C + 8) */
  int64_t ldq, ldkv, ldo;
  int N, T, M, heads, d; /* M <= 32 audio tokens per frame; d % 8 == 0 */
  float softmax_scale;   /* d^-0.5 */
  int dtype;
} mmgt_audio_attention_params;
MMGT_API int mmgt_audio_attention(mmgt_ctx*, const mmgt_audio_attention_params*, void* stream);

/* Temporal self-attention of the motion module: sequences run over the F frames of each (batch, token).
 * qkv: (B*F*T, 3*C) fused projections [q|k|v] of rows ordered (b, f, t); out: (B*F*T, C).
 * Replaces VersatileAttention.forward incl. both rearranges (motion_module.py:351-388). */
MMGT_API int mmgt_temporal_attention(mmgt_ctx*, const void* qkv, void* out, int B, int F, int T, int heads, int d,
                            float scale, int dtype, void* stream);

/* Small element-wise pieces -------------------------------------------------------------------- */
/* Timesteps(dim, flip_sin_to_cos, freq_shift) (diffusers embeddings; unet_3d.py:496): t (B) float32 device. */
MMGT_API int mmgt_timestep_embedding(mmgt_ctx*, const float* t, float* out, int B, int dim, int flip_sin_to_cos,
                            float freq_shift, void* stream);
/* y = silu(x), float32 (resnet.py:226 nonlinearity on temb). */
MMGT_API int mmgt_silu_f32(mmgt_ctx*, const float* x, float* y, int64_t n, void* stream);
/* nearest x2 upsample of (N,H,W,C) -> (N,2H,2W,C) (resnet.py:71-73). */
MMGT_API int mmgt_upsample_nearest2x(mmgt_ctx*, const void* x, void* y, int N, int H, int W, int C, int dtype, void* stream);
/* im2col for 3x3 / pad 1 with stride and optional x2 nearest upsample: (N,H,W,C) -> (N*Ho*Wo, 9*C). */
MMGT_API int mmgt_im2col3x3(mmgt_ctx*, const void* x, void* col, int N, int H, int W, int C, int stride, int upsample2x,
                   int dtype, void* stream);

/* dst (rows, Cpad) = [src (rows, C) | 0]: pads the 4 latent channels so that conv_in (unet_3d.py:517) runs as a
 * tensor-core implicit GEMM (its weight is zero-padded to the same width). */
MMGT_API int mmgt_pad_channels(mmgt_ctx*, const void* src, void* dst, int64_t rows, int C, int Cpad, int dtype, void* stream);

/* dst[i, :] = src[idx[i], :]; rows of row_bytes (multiple of 16) bytes.  Assembles a context window from
 * whole-video tensors (latents[:, :, c], pose_fea[:, :, c], masks.view(2, L, -1)[:, c];
 * pipeline_pose2vid_long.py:556-586). */
MMGT_API int mmgt_gather_rows(mmgt_ctx*, const void* src, const int32_t* idx, void* dst, int n_out, int64_t row_bytes,
                              void* stream);

/* Denoise-loop pieces (pipeline_pose2vid_long.py:622-635 + diffusers DDIMScheduler.step) ---------- */
/* noise_acc[(b), c, frames[j], :, :] += pred[(b), c, j, :, :]; noise_acc float32 (2|1,C,L,H,W);
 * pred (Bp,C,Fw,H,W) in pred_dtype, added into batch rows [b0, b0+Bp). */
MMGT_API int mmgt_window_accumulate(mmgt_ctx*, float* noise_acc, const void* pred, const int32_t* frames, int Bp, int b0,
                           int C, int L, int Fw, int HW, int pred_dtype, void* stream);
/* latents = cx*latents + cv*(u + g*(c-u)), u/c = noise_acc[0|1]/count[frame]; cfg=0 => v = noise_acc[0]/count. */
MMGT_API int mmgt_cfg_ddim_step(mmgt_ctx*, float* latents, const float* noise_acc, const float* inv_count, int C, int L,
                       int HW, int cfg, float guidance, float cx, float cv, void* stream);

/* Motion-mask pyramid (src/dataset/image_processor.py:75-102,311-333) ------------------------------ */
/* src: (L, Hs, Ws) uint8.  Bit-exact Pillow 8-bit bilinear resize (horizontal then vertical, 22-bit fixed
 * point) to (L, S, S), then out = offset + u8/255 in float32 (offset 1.0 builds "1 + lips",
 * scripts/audio2vid.py:475).  tmp: >= L*Hs*S bytes scratch.  out_u8 may be NULL. */
MMGT_API int mmgt_mask_resize(mmgt_ctx*, const uint8_t* src, uint8_t* tmp, uint8_t* out_u8, float* out_f32, int L, int Hs,
                     int Ws, int S, float offset, void* stream);

/* Peer memory over NVLink (one process per GPU) -------------------------------- */
/* The ONLY entry points that allocate: a receive buffer other processes can map (cudaMalloc + CUDA IPC).
 * export/import move the 64-byte IPC handle through host memory (the host exchanges it with torch.distributed). */
MMGT_API int mmgt_peer_alloc(mmgt_ctx*, int64_t bytes, void** out_ptr);
MMGT_API int mmgt_peer_free(mmgt_ctx*, void* ptr);
MMGT_API int mmgt_peer_export(mmgt_ctx*, const void* ptr, unsigned char* handle64_host);
MMGT_API int mmgt_peer_import(mmgt_ctx*, const unsigned char* handle64_host, void** out_ptr);
MMGT_API int mmgt_peer_unmap(mmgt_ctx*, void* ptr);
typedef struct {
  uint32_t* flags[MMGT_MAX_PEERS]; /* flags[s]: the MMGT_MAX_PEERS-slot flag array living on shard s (peer-mapped) */
  uint32_t* epoch;                 /* local device counter, starts at 0 */
  uint32_t* status;                /* local device word: set to 1 if a wait timed out */
  int k, my;
  int timeout_ms;                  /* spin limit per barrier (0 = 2000) */
} mmgt_peer_barrier_params;
/* Stream-ordered barrier over the k shards: every store issued by earlier kernels of this stream (including the
 * peer stores of a row exchange) is visible to all shards once their barrier has passed.  One 32-thread kernel:
 * release-store epoch+1 into slot `my` of every peer's flag array, acquire-spin on the own array.  CUDA-graph safe. */
MMGT_API int mmgt_peer_barrier(mmgt_ctx*, const mmgt_peer_barrier_params*, void* stream);
/* The same row exchange as a stand-alone copy (float32 mode, tests, unfused baseline): src (rows, C) with leading
 * dimension lds -> peers.  rows = B*(F/k)*T. */
MMGT_API int mmgt_row_exchange_copy(mmgt_ctx*, const void* src, int64_t lds, int C, int dtype, const mmgt_row_exchange*,
                                    void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MMGT_B200_H */
