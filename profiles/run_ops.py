#!/usr/bin/env python
"""Launch the hot kernels once each at BASELINE config-2 shapes (one B=2 window: 24 frames of 64x64 latent).

Used under ncu (``--profile-from-start off``: only the region between cudaProfilerStart/Stop is captured):

  ncu --set full --clock-control none --import-source on --profile-from-start off \
      -o gpurun_out/ops python profiles/run_ops.py [op ...]

and standalone to event-time each op (``--time``), L2 flushed between repetitions.
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from mmgt_b200.kernels import get_engine  # noqa: E402
from mmgt_b200.packing import geglu_interleave, ln_fold, subpixel_pack  # noqa: E402


def make_ops(eng, dev):
    bf = torch.bfloat16
    g = torch.Generator(device="cpu").manual_seed(0)

    def rnd(*shape, dtype=bf, scale=1.0):
        return (torch.randn(*shape, generator=g) * scale).to(device=dev, dtype=dtype)

    ops = {}
    N, T0 = 24, 4096
    # ---- spatial attention, level 0 (d = 40), [self ; reference] keys for the cond half
    for name, C, T in (("attn_d40", 320, 4096), ("attn_d80", 640, 1024), ("attn_d160", 1280, 256)):
        qkv = rnd(N, T, 3 * C)
        kv2 = rnd(2, T, 2 * C)
        seg2 = torch.tensor([-1] * 12 + [1] * 12, dtype=torch.int32, device=dev)
        ops[name] = (lambda qkv=qkv, kv2=kv2, seg2=seg2, C=C: eng.attention(
            qkv[:, :, :C], qkv[:, :, C:2 * C], qkv[:, :, 2 * C:], 8, k2=kv2[:, :, :C], v2=kv2[:, :, C:], seg2_index=seg2),
            4.0 * 8 * T * (12 * T + 12 * 2 * T) * (C // 8), 0.0)
    qkv = rnd(N, T0, 960)
    ops["attn_d40_self"] = (lambda: eng.attention(qkv[:, :, :320], qkv[:, :, 320:640], qkv[:, :, 640:], 8),
                            4.0 * 8 * N * T0 * T0 * 40, 0.0)
    # ---- MM-HAA audio cross attention (Lk = 32)
    q3 = rnd(N, T0, 960)
    kv6 = rnd(N, 32, 1920)
    ops["attn_audio_d40"] = (lambda: eng.attention(q3[:, :, :320], kv6[:, :, :320], kv6[:, :, 320:640], 8),
                             4.0 * 8 * N * T0 * 32 * 40, 2.0 * N * T0 * 320 * 2)
    # ---- fused three-region MM-HAA cross attention (mask gate in the epilogue) + its K = 3C+8 GEMM
    masks = [torch.rand(N * T0, generator=g).to(dev) for _ in range(3)]
    q3f, kv6f = q3.view(N * T0, 960), kv6.view(N * 32, 1920)
    ops["audio_attention_fused_d40"] = (lambda: eng.audio_attention(q3f, kv6f, masks, (1.0, 1.0, 2.0), N, T0, 8),
                                        4.0 * 3 * 8 * N * T0 * 32 * 40, (N * T0 * 960 + N * T0 * 968) * 2.0)
    a968 = rnd(N * T0, 968)
    w968 = rnd(320, 968, scale=0.03)
    res968 = rnd(N * T0, 320)
    b968 = rnd(320, dtype=torch.float32)
    ops["gemm_320x968_res"] = (lambda: eng.gemm(a968, w968, bias=b968, residual=res968), 2.0 * N * T0 * 320 * 968,
                               (N * T0 * 968 + 2 * N * T0 * 320) * 2.0)
    # ---- GEMMs (rows = 24 * 4096)
    M = N * T0
    a320 = rnd(M, 320)
    res320 = rnd(M, 320)
    w = rnd(320, 320, scale=0.05)
    b = rnd(320, dtype=torch.float32)
    ops["gemm_320x320_res"] = (lambda: eng.gemm(a320, w, bias=b, residual=res320), 2.0 * M * 320 * 320, 3.0 * M * 320 * 2)
    wq = rnd(960, 320, scale=0.05)
    ops["gemm_960x320"] = (lambda: eng.gemm(a320, wq), 2.0 * M * 960 * 320, (M * 320 + M * 960) * 2.0)
    w1 = torch.randn(2560, 320, generator=g) * 0.05
    b1 = torch.randn(2560, generator=g)
    gb = eng.geglu_block(2560)
    w1i, b1i = geglu_interleave(w1, b1, gb)
    w1i, b1i = w1i.to(dev, bf).contiguous(), b1i.to(dev, torch.float32).contiguous()
    ops["gemm_geglu_2560x320"] = (lambda: eng.gemm(a320, w1i, bias=b1i, geglu_block=gb), 2.0 * M * 2560 * 320,
                                  (M * 320 + M * 1280) * 2.0)
    a1280 = rnd(M, 1280)
    w2 = rnd(320, 1280, scale=0.05)
    ops["gemm_320x1280_res"] = (lambda: eng.gemm(a1280, w2, bias=b, residual=res320), 2.0 * M * 320 * 1280,
                                (M * 1280 + 2 * M * 320) * 2.0)
    M2 = N * 256
    a2 = rnd(M2, 1280)
    w3 = rnd(1280, 1280, scale=0.05)
    b3 = rnd(1280, dtype=torch.float32)
    r3 = rnd(M2, 1280)
    ops["gemm_1280x1280_res_m6144"] = (lambda: eng.gemm(a2, w3, bias=b3, residual=r3), 2.0 * M2 * 1280 * 1280,
                                       (3.0 * M2 * 1280 + 1280 * 1280) * 2)
    # ---- norms
    gam, bet = rnd(320, dtype=torch.float32), rnd(320, dtype=torch.float32)
    ops["layernorm_320"] = (lambda: eng.layernorm(a320, gam, bet), 0.0, 2.0 * M * 320 * 2)
    x4 = rnd(N, 64, 64, 320)
    ops["groupnorm_320_silu"] = (lambda: eng.groupnorm(x4, None, gam, bet, 32, 1e-5, True), 0.0, 2.0 * M * 320 * 2)
    x4b = rnd(N, 16, 16, 1280)
    g2, be2 = rnd(1280, dtype=torch.float32), rnd(1280, dtype=torch.float32)
    ops["groupnorm_1280_silu"] = (lambda: eng.groupnorm(x4b, None, g2, be2, 32, 1e-5, True), 0.0, 2.0 * M2 * 1280 * 2)
    # ---- conv 3x3
    wc = rnd(320, 3, 3, 320, scale=0.02)
    ops["conv3x3_320_320"] = (lambda: eng.conv3x3(x4, wc, bias=b, residual=x4), 2.0 * M * 9 * 320 * 320,
                              3.0 * M * 320 * 2)
    wc2 = rnd(1280, 3, 3, 1280, scale=0.02)
    ops["conv3x3_1280_1280"] = (lambda: eng.conv3x3(x4b, wc2, bias=b3, residual=x4b), 2.0 * M2 * 9 * 1280 * 1280,
                                (3.0 * M2 * 1280 + 9 * 1280 * 1280) * 2)
    # ---- round 2: LayerNorm folded into the consuming GEMM (row statistics + epilogue), strided / upsampling implicit convs
    ops["row_stats_320"] = (lambda: eng.row_stats(a320, 1e-5), 0.0, 1.0 * M * 320 * 2)
    wq_f = torch.randn(960, 320, generator=g) * 0.05
    wq_ln, cs_ln, b_ln = ln_fold(wq_f, None, gam, bet, eng)
    st320 = eng.row_stats(a320, 1e-5)
    ops["gemm_960x320_lnfused"] = (lambda: eng.gemm(a320, wq_ln, bias=b_ln, rowstats=st320, colsum=cs_ln),
                                   2.0 * M * 960 * 320, (M * 320 + M * 960) * 2.0)
    w1_ln, cs1_ln, b1_ln = ln_fold(w1i.float(), b1i, gam, bet, eng)
    ops["gemm_geglu_2560x320_lnfused"] = (lambda: eng.gemm(a320, w1_ln, bias=b1_ln, geglu_block=gb, rowstats=st320, colsum=cs1_ln),
                                          2.0 * M * 2560 * 320, (M * 320 + M * 1280) * 2.0)
    ops["conv3x3_320_320_stride2"] = (lambda: eng.conv3x3(x4, wc, bias=b, stride=2), 2.0 * (M // 4) * 9 * 320 * 320,
                                      (M * 320 + (M // 4) * 320) * 2.0)
    x32 = rnd(N, 32, 32, 640)
    wc640 = torch.randn(640, 640, 3, 3, generator=g) * 0.02
    wc640_k = wc640.permute(0, 2, 3, 1).to(dev, bf).contiguous()
    wc640_sp = subpixel_pack(wc640, eng)
    b640 = rnd(640, dtype=torch.float32)
    ops["conv3x3_640_640_upsample2x"] = (lambda: eng.conv3x3(x32, wc640_k, bias=b640, upsample2x=True, w_subpixel=wc640_sp),
                                         2.0 * M * 9 * 640 * 640, (N * 1024 * 640 + M * 640) * 2.0)
    x4c = rnd(N, 64, 64, 640)
    g640, be640 = rnd(640, dtype=torch.float32), rnd(640, dtype=torch.float32)
    ops["groupnorm_640_silu_64x64"] = (lambda: eng.groupnorm(x4c, None, g640, be640, 32, 1e-5, True), 0.0, 2.0 * M * 640 * 2)
    # ---- temporal attention
    tq = rnd(M, 960)
    ops["temporal_attn_d40"] = (lambda: eng.temporal_attention(tq, 2, 12, T0, 8), 4.0 * 2 * T0 * 8 * 144 * 40,
                                (M * 960 + M * 320) * 2.0)
    return ops


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("ops", nargs="*")
    ap.add_argument("--time", action="store_true")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--attn-v2", action="store_true", help="A/B: the three-S-buffer attention kernel for head dim <= 64")
    ap.add_argument("--attn-q256", type=int, default=None, help="A/B: flag 15")
    ap.add_argument("--attn-packed", type=int, default=None, help="A/B: flag 16")
    ap.add_argument("--ln-persist", type=int, default=None, help="A/B: flag 17")
    ap.add_argument("--attn-persist", action="store_true", help="A/B: head dim <= 64 attention on persistent CTAs (flag 14) instead of one (frame, head, query tile) per CTA")
    ap.add_argument("--lane-residual", action="store_true", help="A/B: residual epilogues with per-lane loads instead of [R | I] k-blocks")
    ap.add_argument("--lane-stores", action="store_true", help="A/B: lean epilogues with per-lane stores instead of TMA stores")
    ap.add_argument("--general-epilogue", action="store_true", help="A/B: GEMM / conv always through the general epilogue code")
    ap.add_argument("--gn-fused", action="store_true", help="A/B: the round-1 single-kernel GroupNorm (spin barrier)")
    ap.add_argument("--conv-im2col", action="store_true", help="A/B: stride-2 / upsampling convs through a staged im2col matrix")
    ap.add_argument("--geglu-exact", action="store_true", help="A/B: erf GELU in the GEGLU epilogue")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    eng = get_engine(dev, torch.bfloat16)
    eng.ctx.set_attention_v2(args.attn_v2)
    eng.ctx.set_attention_persistent(1 if args.attn_persist else 0)
    if args.attn_q256 is not None:
        eng.ctx.set_attention_q256(int(args.attn_q256))
    if args.ln_persist is not None:
        eng.ctx.set_layernorm_persistent(int(args.ln_persist))
    if args.attn_packed is not None:
        eng.ctx.set_attention_packed(bool(args.attn_packed))
    eng.ctx.set_groupnorm_split(not args.gn_fused)
    eng.ctx.set_lean_epilogue(not args.general_epilogue)
    eng.ctx.set_tma_store(not args.lane_stores)
    eng.ctx.set_residual_mma(not args.lane_residual)
    eng.ctx.set_conv_implicit_all(not args.conv_im2col)
    eng.ctx.set_geglu_exact(args.geglu_exact)
    ops = make_ops(eng, dev)
    names = args.ops or list(ops)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    for n in names:          # warm-up (cudaFuncSetAttribute, descriptor paths, allocator)
        for _ in range(2):
            ops[n][0]()
    torch.cuda.synchronize()
    if args.time:
        for n in names:
            fn, flops, nbytes = ops[n]
            ts = []
            for _ in range(args.reps):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fn()
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ts.sort()
            ms = ts[len(ts) // 2]
            print(f"{n:28s} {ms * 1e3:9.1f} us  {flops / ms / 1e9:8.1f} TFLOP/s  {nbytes / ms / 1e6:8.1f} GB/s", flush=True)
        return
    torch.cuda.cudart().cudaProfilerStart()
    for n in names:
        flush.zero_()
        torch.cuda.nvtx.range_push(n)
        ops[n][0]()
        torch.cuda.nvtx.range_pop()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()


if __name__ == "__main__":
    main()
