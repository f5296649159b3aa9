"""Derived, kernel-layout copies of module parameters ("packs").

The state dict stays exactly the reference's (keys, shapes, dtypes) and is owned by
the nn.Module.  Kernels want other layouts (KRSC conv weights, fused q|k|v, GEGLU row interleave,
float32 bias / affine vectors), so each module builds them lazily and rebuilds when a source tensor
changes (``load_state_dict`` bumps ``_version``; ``.to()`` changes ``data_ptr``).
"""
from typing import Callable, Sequence

import torch


class Pack:
    def __init__(self):
        self._key = None
        self._val = None

    def get(self, eng, params: Sequence[torch.Tensor], build: Callable[[], object]):
        key = (str(eng.device), eng.dtype) + tuple((p.data_ptr(), p._version, p.dtype) for p in params)
        if key != self._key:
            with torch.no_grad():
                self._val = build()
            self._key = key
        return self._val


def f32(t: torch.Tensor, eng) -> torch.Tensor:
    return t.detach().to(device=eng.device, dtype=torch.float32).contiguous()


def run(t: torch.Tensor, eng) -> torch.Tensor:
    return t.detach().to(device=eng.device, dtype=eng.dtype).contiguous()


def conv_krsc(weight: torch.Tensor, eng) -> torch.Tensor:
    """(Cout, Cin, 3, 3) -> (Cout, 3, 3, Cin) contiguous in the run dtype."""
    return weight.detach().to(device=eng.device).permute(0, 2, 3, 1).to(eng.dtype).contiguous()


def conv1x1(weight: torch.Tensor, eng) -> torch.Tensor:
    """(Cout, Cin, 1, 1) -> (Cout, Cin)."""
    return run(weight.reshape(weight.shape[0], weight.shape[1]), eng)


def geglu_interleave(weight: torch.Tensor, bias: torch.Tensor, gb: int):
    """diffusers GEGLU.proj rows are [value (n) ; gate (n)].  Re-order to [value(gb) | gate(gb)]* so
    that a tile of 2*gb consecutive rows holds matching value / gate columns."""
    n = weight.shape[0] // 2
    assert n % gb == 0
    wv, wg = weight[:n].reshape(n // gb, gb, -1), weight[n:].reshape(n // gb, gb, -1)
    w = torch.stack([wv, wg], dim=1).reshape(2 * n, -1)
    bv, bg = bias[:n].reshape(n // gb, gb), bias[n:].reshape(n // gb, gb)
    b = torch.stack([bv, bg], dim=1).reshape(2 * n)
    return w, b


def subpixel_upsample_weights(weight: torch.Tensor):
    """Upsample3D = nearest x2 followed by a 3x3 / pad 1 convolution (resnet.py:70-88).  Each output parity (a, b) of
    out[2y+a, 2x+b] only ever sees a 2x2 neighbourhood of the LOW-resolution input, so the operator is four 2x2-tap
    convolutions on the input with pre-summed weights: rows {y-1, y} for a = 0 and {y, y+1} for a = 1 (same for columns),
    zero outside the image -- 16 tap-GEMMs at input resolution instead of 9 at output resolution (2.25x fewer FLOPs) and
    no upsampled / im2col tensor.  ``subpixel_pack`` lays the result out for mmgt_conv3x3 (w_subpixel); the identity is
    checked in tests/test_host_logic.py.

    weight (Cout, Cin, 3, 3) -> dict[(a, b)] = (taps (Cout, Cin, 2, 2), row offsets (2,), column offsets (2,))."""
    w = weight.detach()
    out = {}
    for a in (0, 1):
        # a = 0: upsampled rows 2y-1, 2y, 2y+1 -> input rows y-1, y, y     a = 1: rows 2y, 2y+1, 2y+2 -> y, y, y+1
        rows = ([w[:, :, 0], w[:, :, 1] + w[:, :, 2]], (-1, 0)) if a == 0 else ([w[:, :, 0] + w[:, :, 1], w[:, :, 2]], (0, 1))
        for b in (0, 1):
            taps = []
            for r in rows[0]:                                    # r: (Cout, Cin, 3) over kernel columns
                cols = [r[:, :, 0], r[:, :, 1] + r[:, :, 2]] if b == 0 else [r[:, :, 0] + r[:, :, 1], r[:, :, 2]]
                taps.append(torch.stack(cols, dim=-1))
            out[(a, b)] = (torch.stack(taps, dim=-2), rows[1], (-1, 0) if b == 0 else (0, 1))
    return out


def subpixel_pack(weight: torch.Tensor, eng) -> torch.Tensor:
    """(Cout, Cin, 3, 3) -> (Cout, 4, 2, 2, Cin) in the run dtype: parity index 2a + b, then tap (ty, tx), channels last
    (the ``w_subpixel`` operand of mmgt_conv3x3, include/mmgt_b200.h).  The taps are summed in float32 and rounded once."""
    sp = subpixel_upsample_weights(weight.detach().float())
    per_parity = [sp[(a, b)][0].permute(0, 2, 3, 1) for a in (0, 1) for b in (0, 1)]      # 4 x (Cout, 2, 2, Cin)
    return torch.stack(per_parity, dim=1).to(device=eng.device, dtype=eng.dtype).contiguous()


def ln_fold(weight: torch.Tensor, bias, gamma: torch.Tensor, beta: torch.Tensor, eng):
    """Fold the affine LayerNorm that feeds ``y = LN(x) W^T + bias`` into the GEMM (mmgt_gemm rowstats / colsum):
    LN(x) W^T + bias = rstd * (x W'^T - mean * colsum) + bias'  with  W' = W diag(gamma), colsum = W' 1, bias' = bias + W beta.
    colsum is taken from the ROUNDED W' (what the tensor cores multiply), in float64.  -> (W' run dtype, colsum f32, bias' f32)"""
    w = weight.detach().to(device=eng.device, dtype=torch.float64)
    g = gamma.detach().to(device=eng.device, dtype=torch.float64)
    b = beta.detach().to(device=eng.device, dtype=torch.float64)
    wp = (w * g[None, :]).to(eng.dtype).contiguous()
    colsum = wp.double().sum(dim=1).float().contiguous()
    bp = w @ b
    if bias is not None:
        bp = bp + bias.detach().to(device=eng.device, dtype=torch.float64)
    return wp, colsum, bp.float().contiguous()
