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
