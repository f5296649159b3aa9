"""GPU: every C-ABI operator against the oracle primitives (plain torch fp32 / numpy integer code) on the
same seeded inputs.  Integer work (mask pyramid, gathers) must be bit-exact; float32 kernels 1e-5-ish;
bf16 kernels are compared against the float32 result computed from the same bf16-rounded inputs."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from helpers import rel_l2  # noqa: E402


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    return torch.device("cuda", 0)


def eng_for(dev, dtype):
    from mmgt_b200.kernels import get_engine
    return get_engine(dev, dtype)


def rnd(*shape, dev, dtype=torch.float32, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(device=dev, dtype=dtype)


TOL = {torch.float32: 2e-5, torch.bfloat16: 1.2e-2}


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_layout_roundtrip_and_add(dev, dtype):
    eng = eng_for(dev, dtype)
    x = rnd(2, 5, 3, 6, 7, dev=dev, seed=1)
    a = rnd(2, 5, 3, 6, 7, dev=dev, seed=2)
    tok = eng.ncfhw_to_tokens(x, add=a)
    ref = (x + a).permute(0, 2, 3, 4, 1).reshape(6, 6, 7, 5)
    assert rel_l2(tok.float(), ref) < TOL[dtype]
    back = eng.tokens_to_ncfhw(tok, 2, 3, torch.float32)
    assert torch.equal(back, tok.float().reshape(2, 3, 6, 7, 5).permute(0, 4, 1, 2, 3))


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("shape", [(3, 16, 64, 0), (2, 64, 320, 0), (2, 16, 1280, 640), (4, 256, 128, 64), (2, 9, 1280, 1280),
                                   (24, 4096, 320, 0), (24, 1024, 640, 320), (5, 1000, 320, 0), (300, 7, 64, 0)])
def test_groupnorm_with_virtual_concat_and_silu(dev, dtype, shape):
    N, T, C1, C2 = shape
    eng = eng_for(dev, dtype)
    x1 = rnd(N, T, C1, dev=dev, dtype=dtype, seed=3) * 2 + 0.5
    x2 = rnd(N, T, C2, dev=dev, dtype=dtype, seed=4) if C2 else None
    C = C1 + C2
    g, b = rnd(C, dev=dev, seed=5) * 0.1 + 1, rnd(C, dev=dev, seed=6) * 0.1
    for silu, eps in ((True, 1e-5), (False, 1e-6)):
        y = eng.groupnorm(x1, x2, g, b, 32, eps, silu)
        cat = torch.cat([x1, x2], -1) if C2 else x1
        ref = F.group_norm(cat.float().permute(0, 2, 1), 32, g, b, eps).permute(0, 2, 1)
        ref = F.silu(ref) if silu else ref
        assert y.shape == (N, T, C)
        assert rel_l2(y.float(), ref) < (2e-5 if dtype == torch.float32 else 6e-3)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("C", [64, 320, 640, 1280])
def test_layernorm_and_positional_encoding(dev, dtype, C):
    eng = eng_for(dev, dtype)
    Fr, T = 3, 333
    x = rnd(2 * Fr * T, C, dev=dev, dtype=dtype, seed=7) * 3 + 1
    g, b = rnd(C, dev=dev, seed=8) * 0.1 + 1, rnd(C, dev=dev, seed=9) * 0.1
    pe = rnd(32, C, dev=dev, seed=10)
    ref = F.layer_norm(x.float(), (C,), g, b, 1e-5)
    assert rel_l2(eng.layernorm(x, g, b).float(), ref) < (2e-6 if dtype == torch.float32 else 5e-3)
    frame = (torch.arange(2 * Fr * T, device=dev) // T) % Fr
    assert rel_l2(eng.layernorm(x, g, b, pe=pe, T=T, F=Fr).float(), ref + pe[frame]) < (2e-6 if dtype == torch.float32 else 5e-3)


@pytest.mark.parametrize("C,rows", [(320, 40000), (640, 9001), (1280, 3000), (320, 37)])
def test_layernorm_grid_variants_bit_identical(dev, C, rows):
    """Flag 17: exact-wave persistent LayerNorm grids (every lane group walks several rows, with and without the next-row
    prefetch) must reproduce the default grid bit for bit, with the motion module's positional table too."""
    eng = eng_for(dev, torch.bfloat16)
    x = rnd(rows, C, dev=dev, dtype=torch.bfloat16, seed=17) * 2 + 0.5
    g, b = rnd(C, dev=dev, seed=18) * 0.1 + 1, rnd(C, dev=dev, seed=19) * 0.1
    pe = rnd(24, C, dev=dev, seed=20)
    T = max(1, rows // 24)
    try:
        outs = []
        for mode in (0, 1, 2):
            eng.ctx.set_layernorm_persistent(mode)
            outs.append((eng.layernorm(x, g, b), eng.layernorm(x[:T * 24], g, b, pe=pe, T=T, F=24)))
    finally:
        eng.ctx.set_layernorm_persistent(LN_PERSIST_DEFAULT)
    ref = F.layer_norm(x.float(), (C,), g, b, 1e-5)
    assert rel_l2(outs[0][0].float(), ref) < 5e-3
    for o in outs[1:]:
        assert torch.equal(o[0], outs[0][0]) and torch.equal(o[1], outs[0][1])


def _gemm_ref(A, W, bias, rowscale, rowbias, rpg, residual, alpha, geglu):
    acc = A.float() @ W.float().t()
    if bias is not None:
        acc = acc + bias
    if geglu:
        n = acc.shape[1] // 2
        acc = acc[:, :n] * F.gelu(acc[:, n:])
    if rowscale is not None:
        acc = acc * rowscale[:, None]
    acc = acc * alpha
    if rowbias is not None:
        acc = acc + rowbias[torch.arange(A.shape[0], device=A.device) // rpg]
    if residual is not None:
        acc = acc + residual.float()
    return acc


GEMM_SHAPES = [(2, 1280, 320), (257, 320, 320), (1000, 960, 320), (384, 64, 2880), (130, 160, 72), (96, 4, 64),
               (300, 640, 1280), (128, 256, 64), (513, 192, 64), (64, 1920, 768)]


@pytest.mark.parametrize("dtype,tc", [(torch.float32, False), (torch.bfloat16, False), (torch.bfloat16, True)],
                         ids=["f32", "bf16simt", "bf16tc"])
@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
def test_gemm_full_epilogue(dev, dtype, tc, M, N, K):
    eng = eng_for(dev, dtype)
    eng.ctx.set_tensor_cores(tc)
    try:
        A = rnd(M, K, dev=dev, dtype=dtype, seed=11)
        W = rnd(N, K, dev=dev, dtype=dtype, seed=12, scale=K ** -0.5)
        bias, rs = rnd(N, dev=dev, seed=13), rnd(M, dev=dev, seed=14).abs() + 0.5
        rpg = 37
        rb = rnd((M + rpg - 1) // rpg, N, dev=dev, seed=15)
        res = rnd(M, N, dev=dev, dtype=dtype, seed=16)
        tol = 1e-5 if dtype == torch.float32 else 8e-3
        assert rel_l2(eng.gemm(A, W).float(), _gemm_ref(A, W, None, None, None, 1, None, 1.0, False)) < tol
        out = eng.gemm(A, W, bias=bias, rowscale=rs, rowbias=rb, rows_per_group=rpg, residual=res, alpha=0.7)
        assert rel_l2(out.float(), _gemm_ref(A, W, bias, rs, rb, rpg, res, 0.7, False)) < tol
        # in-place accumulate (residual aliases the output), as the split-K shortcut / MM-HAA sum use it
        acc = res.clone()
        eng.gemm(A, W, bias=bias, residual=acc, out=acc, alpha=2.0)
        assert rel_l2(acc.float(), _gemm_ref(A, W, bias, None, None, 1, res, 2.0, False)) < tol
    finally:
        eng.ctx.set_tensor_cores(True)


@pytest.mark.parametrize("dtype,tc", [(torch.float32, False), (torch.bfloat16, False), (torch.bfloat16, True)],
                         ids=["f32", "bf16simt", "bf16tc"])
@pytest.mark.parametrize("M,C", [(300, 64), (1000, 320), (257, 640)])
def test_gemm_geglu(dev, dtype, tc, M, C):
    from mmgt_b200.packing import geglu_interleave
    eng = eng_for(dev, dtype)
    eng.ctx.set_tensor_cores(tc)
    try:
        A = rnd(M, C, dev=dev, dtype=dtype, seed=21)
        W = rnd(8 * C, C, dev=dev, dtype=dtype, seed=22, scale=C ** -0.5)
        bias = rnd(8 * C, dev=dev, seed=23)
        gb = eng.geglu_block(8 * C)
        Wi, bi = geglu_interleave(W, bias, gb)
        out = eng.gemm(A, Wi.contiguous(), bias=bi.contiguous(), geglu_block=gb)
        assert out.shape == (M, 4 * C)
        assert rel_l2(out.float(), _gemm_ref(A, W, bias, None, None, 1, None, 1.0, True)) < (1e-5 if dtype == torch.float32 else 8e-3)
    finally:
        eng.ctx.set_tensor_cores(True)


# Shapes that take the weight-stationary kernel (B tile resident in shared memory; needs many m-tiles per CTA):
# BN = 160 (N=320), 240 (N=960) and the GEGLU BN = 256; N = 640, K = 640 has no resident plan any more (both runs stream).
# M is ragged on purpose.
BRES_SHAPES = [(40003, 320, 320, False), (38021, 960, 320, False), (15001, 640, 640, False), (37999, 2560, 320, True),
               (40000, 320, 328, False)]


@pytest.mark.parametrize("M,N,K,geglu", BRES_SHAPES)
def test_gemm_weight_stationary_matches_streaming_and_reference(dev, M, N, K, geglu):
    from mmgt_b200.packing import geglu_interleave
    eng = eng_for(dev, torch.bfloat16)
    A = rnd(M, K, dev=dev, dtype=torch.bfloat16, seed=41)
    W = rnd(N, K, dev=dev, dtype=torch.bfloat16, seed=42, scale=K ** -0.5)
    bias = rnd(N, dev=dev, seed=43)
    n_out = N // 2 if geglu else N
    res = None if geglu else rnd(M, n_out, dev=dev, dtype=torch.bfloat16, seed=44)
    rs = None if geglu else rnd(M, dev=dev, seed=45).abs() + 0.5
    rpg = 4096
    rb = None if geglu else rnd((M + rpg - 1) // rpg, N, dev=dev, seed=46)
    gb = 0
    Wk, bk = W, bias
    if geglu:
        gb = eng.geglu_block(N)
        Wk, bk = geglu_interleave(W, bias, gb)
        Wk, bk = Wk.contiguous(), bk.contiguous()
    outs = []
    try:
        for on in (True, False):
            eng.ctx.set_resident_weights(on)
            outs.append(eng.gemm(A, Wk, bias=bk, rowscale=rs, rowbias=rb, rows_per_group=rpg if rb is not None else 0,
                                 residual=res, alpha=0.9, geglu_block=gb))
    finally:
        eng.ctx.set_resident_weights(True)
    ref = _gemm_ref(A, W, bias, rs, rb, rpg, res, 0.9, geglu)
    assert rel_l2(outs[0].float(), ref) < 8e-3
    # same tile shapes are not guaranteed between the two kernels, but both accumulate in fp32 over the same K order
    assert rel_l2(outs[0].float(), outs[1].float()) < 2e-3


def test_gemm_geglu_fast_erf_accuracy(dev):
    """The tensor-core GEGLU epilogue evaluates erf by Abramowitz-Stegun 7.1.26: check it against exact GELU on a
    gate sweep (value = 1, identity-like weights) at bf16 output resolution."""
    from mmgt_b200.packing import geglu_interleave
    eng = eng_for(dev, torch.bfloat16)
    C = 64
    M = 4096
    gate = torch.linspace(-9, 9, M, device=dev)
    A = torch.zeros(M, C, device=dev)
    A[:, 0] = 1.0
    A[:, 1] = gate
    A = A.to(torch.bfloat16)
    W = torch.zeros(8 * C, C, device=dev)
    W[: 4 * C, 0] = 1.0            # value columns = 1
    W[4 * C:, 1] = 1.0             # gate columns = gate
    bias = torch.zeros(8 * C, device=dev)
    gb = eng.geglu_block(8 * C)
    Wi, bi = geglu_interleave(W.to(torch.bfloat16), bias, gb)
    g = A[:, 1].float()
    ref = F.gelu(g)[:, None].expand(M, 4 * C)
    try:
        for exact in (False, True):     # default: logistic-polynomial fit of Phi (<= 8.2e-4 relative); flag 6: erf form
            eng.ctx.set_geglu_exact(exact)
            out = eng.gemm(A, Wi.contiguous(), bias=bi.contiguous(), geglu_block=gb).float()
            err = (out - ref).abs()
            assert float((err / (ref.abs() + 1e-3)).max()) < 6e-3      # bf16 output rounding (2^-8) dominates
            assert float(err.max()) < 2e-2
    finally:
        eng.ctx.set_geglu_exact(False)


def test_gemm_strided_views_and_f32_vectors(dev):
    eng = eng_for(dev, torch.bfloat16)
    buf = rnd(500, 960, dev=dev, dtype=torch.bfloat16, seed=31)
    W = rnd(320, 320, dev=dev, dtype=torch.bfloat16, seed=32, scale=0.05)
    out = eng.gemm(buf[:, 320:640], W)                       # strided A (lda = 960)
    assert rel_l2(out.float(), buf[:, 320:640].float() @ W.float().t()) < 8e-3
    Wbig = rnd(640, 960, dev=dev, dtype=torch.bfloat16, seed=33, scale=0.03)
    out2 = eng.gemm(buf[:, :320].contiguous(), Wbig[:, 640:])  # strided W (ldw = 960): split-K shortcut
    assert rel_l2(out2.float(), buf[:, :320].float() @ Wbig[:, 640:].float().t()) < 8e-3
    v = rnd(2, 1280, dev=dev, seed=34)
    Wf = rnd(320, 1280, dev=dev, seed=35, scale=0.03)
    o = eng.gemm(v, Wf, bias=rnd(320, dev=dev, seed=36), dtype=torch.float32)
    assert o.dtype == torch.float32 and rel_l2(o, v @ Wf.t() + rnd(320, dev=dev, seed=36)) < 1e-5


CONV_CASES = [  # N, H, W, Cin, Cout, stride, upsample
    (3, 16, 16, 64, 64, 1, 0), (2, 8, 8, 128, 256, 1, 0), (5, 4, 4, 256, 256, 1, 0), (2, 32, 32, 64, 128, 1, 0),
    (2, 64, 64, 64, 64, 1, 0), (3, 2, 2, 256, 256, 1, 0), (2, 16, 16, 64, 64, 2, 0), (2, 8, 8, 128, 128, 1, 1),
    (2, 16, 16, 4, 64, 1, 0), (2, 16, 16, 64, 4, 1, 0), (3, 16, 16, 320, 320, 1, 0), (2, 8, 8, 1920, 640, 1, 0),
    # widths no run of 128 consecutive rows covers (config 5: 96 / 48 / 24 / 12 latents): patch tiles pw x ph x frames
    (2, 96, 96, 64, 64, 1, 0), (1, 48, 48, 128, 64, 1, 0), (3, 24, 24, 64, 128, 1, 0), (5, 12, 12, 128, 128, 1, 0),
    (3, 6, 6, 256, 256, 1, 0), (2, 24, 48, 64, 64, 1, 0),
    # Downsample3D / Upsample3D at the model's shapes and at the config-5 widths: implicit GEMM on the tensor-core tier
    # (TMA traversal stride 2; four 2x2-tap sub-pixel convolutions), im2col / CUDA cores otherwise
    (2, 64, 64, 320, 320, 2, 0), (3, 32, 32, 640, 640, 2, 0), (3, 16, 16, 1280, 1280, 2, 0), (2, 8, 8, 1280, 1280, 1, 1),
    (1, 16, 16, 1280, 1280, 1, 1), (2, 32, 32, 640, 640, 1, 1), (1, 96, 96, 64, 64, 2, 0), (2, 48, 48, 64, 128, 2, 0),
    (3, 24, 24, 64, 64, 2, 0), (2, 12, 12, 128, 64, 1, 1), (1, 24, 24, 64, 64, 1, 1), (5, 4, 4, 256, 128, 1, 1),
    (3, 4, 4, 128, 128, 2, 0)]


@pytest.mark.parametrize("dtype,tc", [(torch.float32, False), (torch.bfloat16, False), (torch.bfloat16, True)],
                         ids=["f32", "bf16simt", "bf16tc"])
@pytest.mark.parametrize("case", CONV_CASES)
def test_conv3x3_fused_epilogue(dev, dtype, tc, case):
    N, H, W, Cin, Cout, stride, up = case
    eng = eng_for(dev, dtype)
    eng.ctx.set_tensor_cores(tc)
    try:
        x = rnd(N, H, W, Cin, dev=dev, dtype=dtype, seed=41)
        w = rnd(Cout, Cin, 3, 3, dev=dev, dtype=dtype, seed=42, scale=(9 * Cin) ** -0.5)
        bias = rnd(Cout, dev=dev, seed=43)
        fpg = 1 if N % 2 else 2
        rb = rnd(N // fpg, Cout, dev=dev, seed=44)
        xin = x.float().permute(0, 3, 1, 2)
        if up:
            xin = F.interpolate(xin, scale_factor=2.0, mode="nearest")
        ref = F.conv2d(xin, w.float(), bias, stride=stride, padding=1)
        ref = ref + rb.repeat_interleave(fpg, 0)[:, :, None, None]
        res = rnd(*ref.permute(0, 2, 3, 1).shape, dev=dev, dtype=dtype, seed=45)
        ref = ref.permute(0, 2, 3, 1) + res.float()
        wsp = None
        if up and eng.subpixel_upsample:
            from mmgt_b200.packing import subpixel_pack
            wsp = subpixel_pack(w, eng)
        n_simt = eng.ctx.simt_launches()
        out = eng.conv3x3(x, w.permute(0, 2, 3, 1).contiguous(), bias=bias, rowbias=rb, frames_per_group=fpg, residual=res,
                          stride=stride, upsample2x=bool(up), w_subpixel=wsp)
        assert out.shape == ref.shape
        assert rel_l2(out.float(), ref) < (1e-5 if dtype == torch.float32 else 8e-3)
        if tc and Cin % 64 == 0 and Cout % 32 == 0:
            assert eng.ctx.simt_launches() == n_simt, "a tensor-core shape fell back to the CUDA-core kernel"
            if stride == 2 or up:      # implicit GEMM vs the staged im2col form of the same operator
                eng.ctx.set_conv_implicit_all(False)
                staged = eng.conv3x3(x, w.permute(0, 2, 3, 1).contiguous(), bias=bias, rowbias=rb, frames_per_group=fpg,
                                     residual=res, stride=stride, upsample2x=bool(up))
                eng.ctx.set_conv_implicit_all(True)
                assert rel_l2(out.float(), staged.float()) < 6e-3
    finally:
        eng.ctx.set_tensor_cores(True)
        eng.ctx.set_conv_implicit_all(True)


LEAN_GEMMS = [(40003, 320, 320, False), (38021, 960, 320, False), (37999, 2560, 320, True), (257, 320, 320, False),
              (1000, 960, 320, False), (513, 192, 64, False), (96, 1280, 1280, False), (130, 160, 72, False),
              (3001, 320, 968, False), (6144, 1280, 1280, False), (300, 640, 64, True)]


@pytest.mark.parametrize("M,N,K,geglu", LEAN_GEMMS)
def test_gemm_specialised_epilogue_matches_general_epilogue(dev, M, N, K, geglu):
    """flag 11: the straight-line epilogue instances (bias / per-tile row bias / GEGLU / residual) against the general
    epilogue on the same launch: identical arithmetic except that bias + row bias are pre-added, so equal to ~1 bf16 ulp;
    clipped last m-tile, column-slice destinations, in-place accumulate."""
    from mmgt_b200.packing import geglu_interleave
    eng = eng_for(dev, torch.bfloat16)
    A = rnd(M, K, dev=dev, dtype=torch.bfloat16, seed=51)
    W = rnd(N, K, dev=dev, dtype=torch.bfloat16, seed=52, scale=K ** -0.5)
    bias = rnd(N, dev=dev, seed=53)
    n_out = N // 2 if geglu else N
    gb = 0
    if geglu:
        gb = eng.geglu_block(N)
        W, bias = geglu_interleave(W, bias, gb)
        W, bias = W.contiguous(), bias.contiguous()
    res = None if geglu else rnd(M, n_out, dev=dev, dtype=torch.bfloat16, seed=54)
    rpg = 256                      # a multiple of the 128-row tile: the row bias is staged per tile
    rb = None if geglu else rnd((M + rpg - 1) // rpg, N, dev=dev, seed=55)
    outs = {}
    try:
        for lean in ("mma", True, "lane", False):
            eng.ctx.set_lean_epilogue(bool(lean))
            eng.ctx.set_residual_mma(lean == "mma")  # flag 13: residual through [R | I] k-blocks on the tensor cores
            eng.ctx.set_tma_store(lean is True)      # flag 12: TMA stores from swizzled boxes vs one 32-byte store per lane
            got = []
            buf = torch.full((M + 3, n_out + 32), 7.0, device=dev, dtype=torch.bfloat16)
            out = buf[:M, 16:16 + n_out]
            eng.gemm(A, W, bias=bias, geglu_block=gb, out=out)                      # bias only
            assert float(buf[:, :16].min()) == 7.0 and float(buf[:, 16 + n_out:].min()) == 7.0 and float(buf[M:].min()) == 7.0
            got.append(out.clone())
            if not geglu:
                got.append(eng.gemm(A, W, bias=bias, residual=res))                  # + residual
                got.append(eng.gemm(A, W, bias=bias, rowbias=rb, rows_per_group=rpg, residual=res))
                got.append(eng.gemm(A, W))                                            # nothing at all
                acc = res.clone()
                eng.gemm(A, W, bias=bias, residual=acc, out=acc)                      # in place
                got.append(acc)
            outs[lean] = got
    finally:
        eng.ctx.set_lean_epilogue(True)
        eng.ctx.set_tma_store(True)
        eng.ctx.set_residual_mma(True)
    for a, b in zip(outs[True], outs[False]):
        assert rel_l2(a.float(), b.float()) < 2e-3
    for a, b in zip(outs["mma"], outs[False]):               # (acc + residual) + bias vs (acc + bias) + residual in fp32
        assert rel_l2(a.float(), b.float()) < 2e-3
    for a, b in zip(outs[True], outs["lane"]):               # the store path does not touch the arithmetic
        assert torch.equal(a, b)
    assert torch.equal(outs[True][0], outs[False][0])       # bias only: the same operations in the same order
    ref = A.float() @ W.float().t() + bias
    if not geglu:
        assert rel_l2(outs[True][1].float(), ref + res.float()) < 8e-3


@pytest.mark.parametrize("case", [(2, 64, 64, 320, 320, 1, 0), (3, 32, 32, 640, 320, 1, 0), (2, 64, 64, 320, 320, 2, 0),
                                  (5, 8, 8, 1280, 1280, 1, 0), (3, 16, 16, 128, 64, 1, 0), (2, 32, 32, 640, 640, 1, 1)])
def test_conv_specialised_epilogue_matches_general_epilogue(dev, case):
    N, H, W, Cin, Cout, stride, up = case
    eng = eng_for(dev, torch.bfloat16)
    x = rnd(N, H, W, Cin, dev=dev, dtype=torch.bfloat16, seed=61)
    w = rnd(Cout, 3, 3, Cin, dev=dev, dtype=torch.bfloat16, seed=62, scale=(9 * Cin) ** -0.5)
    bias = rnd(Cout, dev=dev, seed=63)
    rb = rnd(N, Cout, dev=dev, seed=65)
    Ho, Wo = (2 * H, 2 * W) if up else (H // stride, W // stride)
    res = rnd(N, Ho, Wo, Cout, dev=dev, dtype=torch.bfloat16, seed=64)
    wsp = None
    if up:
        from mmgt_b200.packing import subpixel_pack
        wsp = subpixel_pack(w.permute(0, 3, 1, 2).contiguous(), eng)
    outs = {}
    try:
        for lean in ("mma", True, "lane", False):
            eng.ctx.set_lean_epilogue(bool(lean))
            eng.ctx.set_residual_mma(lean == "mma")
            eng.ctx.set_tma_store(lean is True)
            outs[lean] = [eng.conv3x3(x, w, bias=bias, stride=stride, upsample2x=bool(up), w_subpixel=wsp),
                          eng.conv3x3(x, w, bias=bias, rowbias=rb, frames_per_group=1, residual=res, stride=stride,
                                      upsample2x=bool(up), w_subpixel=wsp)]
    finally:
        eng.ctx.set_lean_epilogue(True)
        eng.ctx.set_tma_store(True)
        eng.ctx.set_residual_mma(True)
    assert torch.equal(outs["mma"][0], outs[False][0])
    assert rel_l2(outs["mma"][1].float(), outs[False][1].float()) < 2e-3
    assert torch.equal(outs[True][0], outs[False][0])
    assert torch.equal(outs[True][0], outs["lane"][0]) and torch.equal(outs[True][1], outs["lane"][1])
    assert rel_l2(outs[True][1].float(), outs[False][1].float()) < 2e-3


@pytest.mark.parametrize("dtype,tc", [(torch.float32, False), (torch.bfloat16, True)], ids=["f32", "bf16tc"])
@pytest.mark.parametrize("act", [1, 2])
def test_conv_and_gemm_activation_epilogue_and_rowbias_slices(dev, dtype, tc, act):
    """SiLU / ReLU epilogue (pose_guider.py:47-57, audio_proj.py:96) and row-bias taken as a column slice of a wider
    matrix (one time-embedding projection for all resnets)."""
    eng = eng_for(dev, dtype)
    eng.ctx.set_tensor_cores(tc)
    fn = F.silu if act == 1 else F.relu
    tol = 1e-5 if dtype == torch.float32 else 8e-3
    try:
        N, H, W, Cin, Cout = 4, 16, 16, 64, 128
        x = rnd(N, H, W, Cin, dev=dev, dtype=dtype, seed=1)
        w = rnd(Cout, Cin, 3, 3, dev=dev, dtype=dtype, seed=2, scale=(9 * Cin) ** -0.5)
        bias = rnd(Cout, dev=dev, seed=3)
        wide = rnd(2, 3 * Cout + 64, dev=dev, seed=4)                # (B, total) float32: this conv owns columns [64, 64 + Cout)
        rb = wide[:, 64:64 + Cout]
        res = rnd(N, H, W, Cout, dev=dev, dtype=dtype, seed=5)
        ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), bias, padding=1) + rb.repeat_interleave(2, 0)[:, :, None, None]
        ref = fn(ref).permute(0, 2, 3, 1) + res.float()
        out = eng.conv3x3(x, w.permute(0, 2, 3, 1).contiguous(), bias=bias, rowbias=rb, frames_per_group=2, residual=res, act=act)
        assert rel_l2(out.float(), ref) < tol
        M, Nn, K = 1000, 256, 192
        A = rnd(M, K, dev=dev, dtype=dtype, seed=6)
        Wg = rnd(Nn, K, dev=dev, dtype=dtype, seed=7, scale=K ** -0.5)
        bg = rnd(Nn, dev=dev, seed=8)
        wide2 = rnd(5, Nn + 32, dev=dev, seed=9)
        rb2 = wide2[:, 32:]
        r2 = rnd(M, Nn, dev=dev, dtype=dtype, seed=10)
        grp = (torch.arange(M, device=dev) // 100) % 5
        refg = fn(A.float() @ Wg.float().t() + bg + rb2[grp]) + r2.float()
        outg = eng.gemm(A, Wg, bias=bg, rowbias=rb2, rows_per_group=100, rowbias_mod=5, residual=r2, act=act)
        assert rel_l2(outg.float(), refg) < tol
    finally:
        eng.ctx.set_tensor_cores(True)


@pytest.mark.parametrize("dtype,tc", [(torch.float32, False), (torch.bfloat16, True)], ids=["f32", "bf16tc"])
@pytest.mark.parametrize("C,N", [(320, 960), (640, 1920), (1280, 3840), (64, 192), (320, 2560)])
def test_layernorm_folded_into_gemm(dev, dtype, tc, C, N):
    """row_stats + GEMM epilogue (rowstats / colsum, packing.ln_fold) == LayerNorm -> Linear, also with the motion module's
    positional table as a per-frame row bias and through the GEGLU epilogue (attention.py:331-362, motion_module.py:365)."""
    from mmgt_b200.packing import geglu_interleave, ln_fold
    eng = eng_for(dev, dtype)
    eng.ctx.set_tensor_cores(tc)
    try:
        Fr, T = 3, 211
        M = 2 * Fr * T
        x = rnd(M, C, dev=dev, dtype=dtype, seed=7) * 2 + 0.7                  # non-zero row means
        g, b = rnd(C, dev=dev, seed=8) * 0.2 + 1, rnd(C, dev=dev, seed=9) * 0.2
        W = rnd(N, C, dev=dev, seed=10, scale=C ** -0.5)
        bias = rnd(N, dev=dev, seed=11)
        ln = F.layer_norm(x.float(), (C,), g, b, 1e-5)
        st = eng.row_stats(x, 1e-5)
        assert rel_l2(st[:, 0], x.float().mean(1)) < 1e-5 and rel_l2(st[:, 1], torch.rsqrt(x.float().var(1, unbiased=False) + 1e-5)) < 1e-4
        wp, colsum, bp = ln_fold(W, bias, g, b, eng)
        tol = 2e-5 if dtype == torch.float32 else 8e-3
        out = eng.gemm(x, wp, bias=bp, rowstats=st, colsum=colsum)
        ref = ln @ W.to(dtype).float().t() + bias
        assert rel_l2(out.float(), ref) < tol
        # + positional table through the projection, added per frame (rows ordered (b, f, t))
        pe = rnd(32, C, dev=dev, seed=12)
        tab = (pe.double() @ W.double().t()).float().contiguous()
        frame = (torch.arange(M, device=dev) // T) % Fr
        out = eng.gemm(x, wp, bias=bp, rowstats=st, colsum=colsum, rowbias=tab[:Fr], rows_per_group=T, rowbias_mod=Fr)
        assert rel_l2(out.float(), (ln + pe[frame]) @ W.to(dtype).float().t() + bias) < tol
        if N % 32 == 0:
            gb = eng.geglu_block(N)
            Wi, bi = geglu_interleave(W, bias, gb)
            wpi, csi, bpi = ln_fold(Wi.contiguous(), bi.contiguous(), g, b, eng)
            outg = eng.gemm(x, wpi, bias=bpi, geglu_block=gb, rowstats=st, colsum=csi)
            full = ln @ W.to(dtype).float().t() + bias
            refg = full[:, : N // 2] * F.gelu(full[:, N // 2:])
            assert rel_l2(outg.float(), refg) < (5e-5 if dtype == torch.float32 else 1e-2)
    finally:
        eng.ctx.set_tensor_cores(True)


def test_strict_tensor_core_mode_and_simt_counter(dev):
    """A bf16 shape no tcgen05 kernel covers runs on CUDA cores and is COUNTED; in strict mode it is an error."""
    from mmgt_b200._lib import MmgtError
    eng = eng_for(dev, torch.bfloat16)
    A = rnd(300, 72, dev=dev, dtype=torch.bfloat16, seed=1)
    W = rnd(100, 72, dev=dev, dtype=torch.bfloat16, seed=2)            # N = 100: no tensor-core tile width divides it
    n0 = eng.ctx.simt_launches()
    out = eng.gemm(A, W)
    assert eng.ctx.simt_launches() == n0 + 1
    assert rel_l2(out.float(), A.float() @ W.float().t()) < 8e-3
    ok = eng.gemm(rnd(300, 64, dev=dev, dtype=torch.bfloat16, seed=3), rnd(128, 64, dev=dev, dtype=torch.bfloat16, seed=4))
    assert eng.ctx.simt_launches() == n0 + 1 and ok.shape == (300, 128)
    eng.ctx.set_strict_tensor_cores(True)
    try:
        with pytest.raises(MmgtError, match="strict"):
            eng.gemm(A, W)
        x = rnd(2, 8, 8, 4, dev=dev, dtype=torch.bfloat16, seed=5)     # Cin = 4: not a tensor-core convolution
        with pytest.raises(MmgtError, match="strict"):
            eng.conv3x3(x, rnd(64, 3, 3, 4, dev=dev, dtype=torch.bfloat16, seed=6))
        q = rnd(2, 16, 64, dev=dev, dtype=torch.bfloat16, seed=7)      # Lq = 16 < 64: CUDA-core attention
        with pytest.raises(MmgtError, match="strict"):
            eng.attention(q, q, q, 8)
        v = rnd(2, 1280, dev=dev, seed=8)                              # batch-sized float32 vectors are CUDA-core by design
        assert eng.gemm(v, rnd(320, 1280, dev=dev, seed=9), dtype=torch.float32).shape == (2, 320)
    finally:
        eng.ctx.set_strict_tensor_cores(False)


def test_conv_in_padded_to_tensor_core_width(dev):
    """InflatedConv3d(4 -> C): input channels zero-padded to 64 so conv_in runs as a tensor-core implicit GEMM."""
    from mmgt_b200.resnet import InflatedConv3d
    eng = eng_for(dev, torch.bfloat16)
    conv = InflatedConv3d(4, 320, kernel_size=3, padding=(1, 1)).to(dev)
    x = rnd(6, 32, 32, 4, dev=dev, dtype=torch.bfloat16, seed=1)
    pose = rnd(6, 32, 32, 320, dev=dev, dtype=torch.bfloat16, seed=2)
    n0 = eng.ctx.simt_launches()
    out = conv.run(eng, x, residual=pose)
    assert eng.ctx.simt_launches() == n0
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), conv.weight.to(torch.bfloat16).float(), conv.bias, padding=1)
    assert rel_l2(out.float(), ref.permute(0, 2, 3, 1) + pose.float()) < 8e-3


@pytest.mark.parametrize("shape", [(24, 4096, 320, 0), (24, 1024, 640, 0), (24, 256, 1280, 1280), (24, 64, 1280, 0),
                                   (3, 4096, 640, 320), (5, 9216, 320, 0)])
def test_groupnorm_split_kernels_match_fused_kernel(dev, shape):
    """The default two-kernel GroupNorm (statistics, normalise) against the single spin-barrier kernel and torch, at the
    model's shapes (config 2 levels 0-3, a skip concat, a 96x96 config-5 frame)."""
    N, T, C1, C2 = shape
    eng = eng_for(dev, torch.bfloat16)
    x1 = rnd(N, T, C1, dev=dev, dtype=torch.bfloat16, seed=1) * 1.5 + 0.3
    x2 = rnd(N, T, C2, dev=dev, dtype=torch.bfloat16, seed=2) if C2 else None
    C = C1 + C2
    g, b = rnd(C, dev=dev, seed=3) * 0.1 + 1, rnd(C, dev=dev, seed=4) * 0.1
    xf = x1.float() if x2 is None else torch.cat([x1.float(), x2.float()], dim=-1)
    ref = F.silu(F.group_norm(xf.transpose(1, 2), 32, g, b, 1e-5).transpose(1, 2))
    outs = []
    try:
        for split in (True, False):
            eng.ctx.set_groupnorm_split(split)
            outs.append(eng.groupnorm(x1, x2, g, b, 32, 1e-5, True))
    finally:
        eng.ctx.set_groupnorm_split(True)
    assert rel_l2(outs[0].float(), ref) < 5e-3 and rel_l2(outs[1].float(), ref) < 5e-3
    assert rel_l2(outs[0].float(), outs[1].float()) < 3e-3


def _attn_ref(q, k, v, heads):
    n, lq, c = q.shape
    d = c // heads
    qh, kh, vh = (t.float().view(n, -1, heads, d).transpose(1, 2) for t in (q, k, v))
    s = (qh @ kh.transpose(-1, -2)) * d ** -0.5
    return (s.softmax(-1) @ vh).transpose(1, 2).reshape(n, lq, c)


LN_PERSIST_DEFAULT = 0        # mmgt_ctx_flag(17) default (csrc/ctx.cu)
ATTN_Q256_DEFAULT = 5       # mmgt_ctx_flag(15) default (csrc/ctx.cu)
ATTN_CASES = [(4, 64, 64, 64, 8, 40), (3, 100, 100, 100, 8, 8), (2, 256, 256, 0, 8, 80), (2, 70, 32, 0, 8, 40),
              (2, 64, 64, 64, 8, 160), (6, 16, 16, 16, 8, 32), (3, 1024, 1024, 1024, 8, 40), (2, 4096, 4096, 4096, 8, 40),
              (2, 300, 300, 300, 8, 80), (2, 256, 256, 256, 8, 160), (2, 1024, 32, 0, 8, 80), (3, 200, 136, 72, 8, 16),
              # ragged shapes: partial q tile, partial key tiles in both segments, several key tiles per softmax group
              (3, 300, 700, 520, 8, 40), (2, 640, 1000, 0, 8, 64), (2, 128, 512, 0, 8, 16), (4, 200, 130, 450, 8, 48)]


@pytest.mark.parametrize("dtype,tc", [(torch.float32, False), (torch.bfloat16, False), (torch.bfloat16, True)],
                         ids=["f32", "bf16simt", "bf16tc"])
@pytest.mark.parametrize("N,Lq,Lk,Lk2,heads,d", ATTN_CASES)
def test_attention_two_segments(dev, dtype, tc, N, Lq, Lk, Lk2, heads, d):
    eng = eng_for(dev, dtype)
    eng.ctx.set_tensor_cores(tc)
    try:
        C = heads * d
        q = rnd(N, Lq, 3 * C, dev=dev, dtype=dtype, seed=51)[:, :, :C]     # frames are Lq rows apart (row stride 3C)
        # keys / values of a frame must be Lk consecutive rows (row stride 3C), frames Lk rows apart
        kvbuf = rnd(N, Lk, 3 * C, dev=dev, dtype=dtype, seed=53)
        k, v = kvbuf[:, :, C:2 * C], kvbuf[:, :, 2 * C:]
        if Lk2:
            bank = rnd(2, Lk2, 2 * C, dev=dev, dtype=dtype, seed=52)
            k2, v2 = bank[:, :, :C], bank[:, :, C:]
            idx = torch.tensor([(-1 if i % 3 == 0 else i % 2) for i in range(N)], dtype=torch.int32, device=dev)
            out = eng.attention(q, k, v, heads, k2=k2, v2=v2, seg2_index=idx)
            if tc and d <= 64:      # the three-S-buffer kernel (flag 9) must agree with the default two-buffer one
                eng.ctx.set_attention_v2(True)
                alt = eng.attention(q, k, v, heads, k2=k2, v2=v2, seg2_index=idx)
                eng.ctx.set_attention_v2(False)
                assert rel_l2(out.float(), alt.float()) < 4e-3
                # persistent CTAs (flag 14) walking many items each -- 3 and 7 CTAs for all (frame, head, query tile) items,
                # frames with one and with two key segments mixed -- vs one item per CTA: same arithmetic, same bits
                eng.ctx.set_attention_persistent(0)
                eng.ctx.set_attention_q256(False)        # the persistent kernel walks 128-query items
                one = eng.attention(q, k, v, heads, k2=k2, v2=v2, seg2_index=idx)
                for ctas in (3, 7):
                    eng.ctx.set_attention_persistent(ctas)
                    per = eng.attention(q, k, v, heads, k2=k2, v2=v2, seg2_index=idx)
                    assert torch.equal(per, one), ctas
                eng.ctx.set_attention_persistent(0)
                eng.ctx.set_attention_q256(ATTN_Q256_DEFAULT)
            refs = []
            for n in range(N):
                kk, vv = k[n:n + 1], v[n:n + 1]
                if idx[n] >= 0:
                    kk = torch.cat([kk, k2[idx[n]:idx[n] + 1]], 1)
                    vv = torch.cat([vv, v2[idx[n]:idx[n] + 1]], 1)
                refs.append(_attn_ref(q[n:n + 1], kk, vv, heads))
            ref = torch.cat(refs)
        else:
            out = eng.attention(q, k, v, heads)
            ref = _attn_ref(q, k, v, heads)
            if tc and d <= 64:
                eng.ctx.set_attention_persistent(0)
                eng.ctx.set_attention_q256(False)
                one = eng.attention(q, k, v, heads)
                eng.ctx.set_attention_persistent(5)
                assert torch.equal(eng.attention(q, k, v, heads), one)
                eng.ctx.set_attention_persistent(0)
                eng.ctx.set_attention_q256(ATTN_Q256_DEFAULT)
        assert rel_l2(out.float(), ref) < (2e-5 if dtype == torch.float32 else 8e-3)
    finally:
        eng.ctx.set_tensor_cores(True)
        eng.ctx.set_attention_v2(False)
        eng.ctx.set_attention_persistent(0)
        eng.ctx.set_attention_q256(ATTN_Q256_DEFAULT)


@pytest.mark.parametrize("N,Lq,Lk,Lk2,heads,d", ATTN_CASES + [(2, 257, 129, 384, 8, 40), (5, 512, 512, 512, 8, 40), (1, 2304, 2304, 0, 8, 40),
                                                        (2, 9216, 9216, 9216, 8, 40)])   # last: the 96 x 96 level of config 5
def test_attention_kernel_variants(dev, N, Lq, Lk, Lk2, heads, d):
    """Flags 15 / 16 of the tensor-core attention: packed fp32 pairs (FFMA2 / FADD2) must not change a bit of either kernel;
    the 256-query kernel (one query tile + one MMA-issuing warp per softmax group, no split-KV merge) and its variants (FMA-pipe
    exp2 for 1 / 2 of 4 score pairs, row sums from the tensor cores) against float32."""
    if Lq < 64:
        pytest.skip("tensor-core attention needs Lq >= 64")
    eng = eng_for(dev, torch.bfloat16)
    C = heads * d
    q = rnd(N, Lq, 3 * C, dev=dev, dtype=torch.bfloat16, seed=71)[:, :, :C]
    kvbuf = rnd(N, Lk, 3 * C, dev=dev, dtype=torch.bfloat16, seed=73)
    k, v = kvbuf[:, :, C:2 * C], kvbuf[:, :, 2 * C:]
    kw = {}
    if Lk2:
        bank = rnd(2, Lk2, 2 * C, dev=dev, dtype=torch.bfloat16, seed=72)
        idx = torch.tensor([(-1 if i % 3 == 0 else i % 2) for i in range(N)], dtype=torch.int32, device=dev)
        kw = dict(k2=bank[:, :, :C], v2=bank[:, :, C:], seg2_index=idx)
    refs = []
    for n in range(N):
        kk, vv = k[n:n + 1], v[n:n + 1]
        if Lk2 and kw["seg2_index"][n] >= 0:
            i = int(kw["seg2_index"][n])
            kk, vv = torch.cat([kk, kw["k2"][i:i + 1]], 1), torch.cat([vv, kw["v2"][i:i + 1]], 1)
        refs.append(_attn_ref(q[n:n + 1], kk, vv, heads))
    ref = torch.cat(refs)
    outs = {}
    try:
        for q256, packed in ((0, False), (0, True), (1, True), (4, True), (5, True), (7, True), (8, True), (9, True)):
            eng.ctx.set_attention_q256(q256)
            eng.ctx.set_attention_packed(packed)
            outs[q256, packed] = o = eng.attention(q, k, v, heads, **kw)
            assert torch.isfinite(o).all(), (q256, packed)
            assert rel_l2(o.float(), ref) < 8e-3, (q256, packed)
    finally:
        eng.ctx.set_attention_q256(ATTN_Q256_DEFAULT)
        eng.ctx.set_attention_packed(True)
    assert torch.equal(outs[0, False], outs[0, True])         # packed pairs: same IEEE operations
    if d <= 64 and Lq > 128:
        # FMA-pipe exp2 (7.5e-5 relative) vs MUFU, and row sums from the tensor cores (sum of the bf16 weights) vs fp32 sums
        # in registers: both far inside the bf16 rounding of P
        for v in (4, 5, 7, 8, 9):
            assert rel_l2(outs[v, True].float(), outs[1, True].float()) < 3e-3, v
    if d > 48:                                                 # no free accumulator columns: the register-sum kernels run
        assert torch.equal(outs[7, True], outs[5, True]) and torch.equal(outs[8, True], outs[1, True])


def test_attention_running_max_jumps_late(dev):
    """Scores that grow along the key axis (every tile's maximum exceeds the previous reference by far more than the lazy
    rescale threshold) and a huge outlier in the LAST key tile: exercises the redo-the-tile path and the O rescale of both
    head-dim <= 64 kernels, plus a partial last tile in each segment."""
    eng = eng_for(dev, torch.bfloat16)
    N, Lq, Lk, Lk2, heads, d = 2, 192, 700, 333, 2, 40
    C = heads * d
    g = torch.Generator().manual_seed(3)
    q = torch.randn(N, Lq, C, generator=g)
    k = torch.randn(N, Lk, C, generator=g) * torch.linspace(0.2, 6.0, Lk)[None, :, None]
    v = torch.randn(N, Lk, C, generator=g)
    k2 = torch.randn(1, Lk2, C, generator=g) * 0.3
    k2[:, -5] = 9.0 * q[0, 7].sign()                   # one key of the last (partial) tile dominates for some rows
    v2 = torch.randn(1, Lk2, C, generator=g)
    q, k, v, k2, v2 = (t.to(device=dev, dtype=torch.bfloat16) for t in (q, k, v, k2, v2))
    ref = _attn_ref(q, torch.cat([k, k2.expand(N, -1, -1)], 1), torch.cat([v, v2.expand(N, -1, -1)], 1), heads)
    try:
        for v2_kernel, persist, q256 in ((True, 0, 0), (False, 0, 0), (False, 2, 0), (False, 0, 1), (False, 0, 5), (False, 0, 7), (False, 0, 9)):
            eng.ctx.set_attention_v2(v2_kernel)
            eng.ctx.set_attention_persistent(persist)
            eng.ctx.set_attention_q256(q256)
            eng.ctx.set_attention_packed(bool(q256))
            out = eng.attention(q, k, v, heads, k2=k2, v2=v2)
            assert torch.isfinite(out).all()
            assert rel_l2(out.float(), ref) < 1e-2, (v2_kernel, persist, q256)
    finally:
        eng.ctx.set_attention_v2(False)
        eng.ctx.set_attention_persistent(0)
        eng.ctx.set_attention_q256(ATTN_Q256_DEFAULT)
        eng.ctx.set_attention_packed(True)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B,Fr,T,heads,d", [(2, 12, 16, 8, 40), (1, 8, 64, 8, 8), (2, 4, 9, 8, 160), (1, 32, 4, 8, 80),
                                            (2, 12, 1024, 8, 80), (1, 16, 333, 8, 40), (3, 5, 7, 8, 64), (2, 12, 4096, 8, 40),
                                            (1, 3, 50, 4, 16), (24, 12, 64, 8, 160)])
def test_temporal_attention(dev, dtype, B, Fr, T, heads, d):
    eng = eng_for(dev, dtype)
    C = heads * d
    qkv = rnd(B * Fr * T, 3 * C, dev=dev, dtype=dtype, seed=61)
    out = eng.temporal_attention(qkv, B, Fr, T, heads)
    x = qkv.float().view(B, Fr, T, 3 * C).permute(0, 2, 1, 3).reshape(B * T, Fr, 3 * C)   # (b d) f c
    ref = _attn_ref(x[:, :, :C], x[:, :, C:2 * C], x[:, :, 2 * C:], heads)
    ref = ref.view(B, T, Fr, C).permute(0, 2, 1, 3).reshape(B * Fr * T, C)
    assert rel_l2(out.float(), ref) < (2e-5 if dtype == torch.float32 else 8e-3)
    if dtype == torch.bfloat16:     # row-coalesced kernel (default for head dim <= 80) vs one warp per (batch, pixel, head)
        try:
            eng.ctx.set_temporal_rows(False)
            old = eng.temporal_attention(qkv, B, Fr, T, heads)
        finally:
            eng.ctx.set_temporal_rows(True)
        assert rel_l2(out.float(), old.float()) < 1e-3


def test_mask_pyramid_bit_exact(dev):
    from mmgt_b200.image_processor import MaskPyramid
    from oracle.mask_pyramid import full_mask_from_lips, mask_pyramid_u8, preprocess_mov_mask
    from oracle.synthetic import synthetic_masks_u8
    import os
    from helpers import GOLD
    real = np.load(os.path.join(GOLD, "real_masks.npz"))      # the reference's bundled case, oracle/make_golden_masks.py
    for face, lips in (synthetic_masks_u8(7), (real["face"], real["lips"])):
        for image_size in (512, 256, 768):
            mp = MaskPyramid(image_size, dev)
            f_gpu, l_gpu = mp.preprocess_mov_mask(list(face), list(lips))
            f_ref, l_ref = preprocess_mov_mask(face, lips, image_size)
            full_gpu, full_ref = mp.full_mask_from_lips(list(lips)), full_mask_from_lips(l_ref)
            for k in range(4):
                assert torch.equal(f_gpu[k].cpu(), torch.from_numpy(f_ref[k])), (image_size, k)
                assert torch.equal(l_gpu[k].cpu(), torch.from_numpy(l_ref[k])), (image_size, k)
                assert torch.equal(full_gpu[k].cpu(), torch.from_numpy(full_ref[k])), (image_size, k)
    noise = np.random.default_rng(5).integers(0, 256, (3, 64, 64), dtype=np.uint8)
    eng = eng_for(dev, torch.float32)
    for s, ref in zip((64, 32, 16, 8), mask_pyramid_u8(noise, 512)):
        _, u8 = eng.mask_resize(torch.from_numpy(noise).to(dev), s, want_u8=True)
        assert torch.equal(u8.cpu(), torch.from_numpy(ref))


def test_small_pieces(dev):
    eng = eng_for(dev, torch.float32)
    t = torch.tensor([500.0, 999.0, 32.0], device=dev)
    emb = eng.timestep_embedding(t, 320, True, 0.0)
    half = 160
    freq = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=dev) / half)
    arg = t[:, None] * freq[None]
    assert torch.allclose(emb, torch.cat([arg.cos(), arg.sin()], -1), atol=2e-4)
    x = rnd(3, 1280, dev=dev, seed=71)
    assert torch.allclose(eng.silu_f32(x), F.silu(x), atol=1e-6)
    src = rnd(10, 24, dev=dev, seed=72)
    idx = torch.tensor([9, 0, 3, 3, 7], dtype=torch.int32, device=dev)
    assert torch.equal(eng.gather_rows(src, idx), src[idx.long()])
    up = rnd(2, 3, 5, 8, dev=dev, seed=73)
    ref = F.interpolate(up.permute(0, 3, 1, 2), scale_factor=2.0, mode="nearest").permute(0, 2, 3, 1)
    assert torch.equal(eng.upsample_nearest2x(up), ref)
    # window accumulate + CFG / DDIM update
    L, Fw = 9, 4
    acc = torch.zeros(2, 4, L, 3, 3, device=dev)
    pred = rnd(1, 4, Fw, 3, 3, dev=dev, seed=74)
    frames = torch.tensor([7, 8, 0, 1], dtype=torch.int32, device=dev)
    eng.window_accumulate(acc, pred, frames, 1)
    eng.window_accumulate(acc, pred, frames, 1)
    ref_acc = torch.zeros_like(acc)
    ref_acc[1, :, frames.long()] = 2 * pred[0]
    assert torch.allclose(acc, ref_acc)
    lat = rnd(1, 4, L, 3, 3, dev=dev, seed=75)
    noise = rnd(2, 4, L, 3, 3, dev=dev, seed=76)
    inv = 1.0 / torch.tensor([1, 2, 1, 1, 2, 2, 1, 1, 2], dtype=torch.float32, device=dev)
    u, c = (noise * inv.view(1, 1, L, 1, 1)).chunk(2)
    want = 0.9 * lat + (-0.3) * (u + 3.5 * (c - u))
    eng.cfg_ddim_step(lat, noise, inv, True, 3.5, 0.9, -0.3)
    assert torch.allclose(lat, want, atol=1e-5)


@pytest.mark.parametrize("N,T,M,heads,d", [(3, 100, 32, 8, 40), (2, 256, 32, 8, 8), (2, 64, 20, 8, 80), (1, 16, 32, 4, 160),
                                           (2, 300, 32, 8, 16), (2, 1024, 32, 8, 40)])
def test_audio_attention_three_regions_gated(dev, N, T, M, heads, d):
    """mmgt_audio_attention: the three MM-HAA cross-attentions (attention.py:719-750) with mask gate x motion_scale in
    the epilogue, plus the gate columns, against softmax(q k^T) v * gate in float32 from the same bf16 inputs."""
    eng = eng_for(dev, torch.bfloat16)
    C = heads * d
    q3 = rnd(N * T, 3 * C, dev=dev, dtype=torch.bfloat16, seed=31)
    kv6 = rnd(N * M, 6 * C, dev=dev, dtype=torch.bfloat16, seed=32)
    masks = [torch.rand(N * T, generator=torch.Generator().manual_seed(40 + r)).to(dev) + (1.0 if r == 0 else 0.0) for r in range(3)]
    scale = (1.0, 1.5, 2.0)
    out = eng.audio_attention(q3, kv6, masks, scale, N, T, heads).float()
    assert out.shape == (N * T, 3 * C + 8)
    for r in range(3):
        q = q3[:, r * C:(r + 1) * C].float().view(N, T, heads, d).transpose(1, 2)
        k = kv6[:, 2 * r * C:(2 * r + 1) * C].float().view(N, M, heads, d).transpose(1, 2)
        v = kv6[:, (2 * r + 1) * C:(2 * r + 2) * C].float().view(N, M, heads, d).transpose(1, 2)
        o = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(N * T, C)
        gate = masks[r] * scale[r]
        assert rel_l2(out[:, r * C:(r + 1) * C], o * gate[:, None]) < TOL[torch.bfloat16]
        assert torch.equal(out[:, 3 * C + r], gate.to(torch.bfloat16).float())
    assert torch.count_nonzero(out[:, 3 * C + 3:]) == 0


def test_mmhaa_fused_regions_match_per_region_operators(dev):
    """AudioTemporalBasicTransformerBlock: fused three-region kernel + one K = 3C+8 GEMM vs the per-region chain
    (attention -> to_out * mask -> zero_conv * scale -> sum), both bf16, against each other and a float32 run."""
    from mmgt_b200.attention import AudioTemporalBasicTransformerBlock
    torch.manual_seed(5)
    C, heads, N, T, M = 320, 8, 4, 256, 32
    blk = AudioTemporalBasicTransformerBlock(C, heads, C // heads, cross_attention_dim=768, depth=0, unet_block_name="down",
                                             stack_enable_blocks_name=["down"], stack_enable_blocks_depth=[0]).to(dev)
    with torch.no_grad():
        for z in (blk.zero_conv_full, blk.zero_conv_face, blk.zero_conv_lip):
            z.weight.normal_(0, 0.05)
            z.bias.normal_(0, 0.05)
    x = rnd(N, T, C, dev=dev, seed=51)
    audio = rnd(N * M, 768, dev=dev, seed=52)
    masks = [torch.rand(N * T, generator=torch.Generator().manual_seed(60 + r)).to(dev) for r in range(3)]
    scale = (1.0, 1.0, 2.0)
    e32, e16 = eng_for(dev, torch.float32), eng_for(dev, torch.bfloat16)
    ref = blk.run(e32, x, audio, masks, scale).float()
    xb, ab = x.to(torch.bfloat16), audio.to(torch.bfloat16)
    blk.fuse_regions = True
    fused = blk.run(e16, xb, ab, masks, scale).float()
    blk.fuse_regions = False
    chain = blk.run(e16, xb, ab, masks, scale).float()
    e_f, e_c = rel_l2(fused, ref), rel_l2(chain, ref)
    print(f"MM-HAA block bf16 vs float32: fused {e_f:.3e}, per-region {e_c:.3e}")
    assert e_f < 1.2e-2 and e_c < 1.2e-2
    assert e_f < 1.5 * e_c + 1e-3
