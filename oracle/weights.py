"""Deterministic, construction-order-independent synthetic weights.  TEST INFRASTRUCTURE.

Every tensor is drawn from its own CPU generator seeded by a hash of (seed, key), so the
reference modules (via ``load_state_dict``), the oracle restatement and the CUDA model all see
bit-identical parameters regardless of how/when their modules were constructed.

Scales follow SURVEY.md section 8d: default-init-like uniform(+-1/sqrt(fan_in)) for weights, and
N(0, 0.02^2) for the tensors the reference zero-initialises (``zero_conv_*``,
``temporal_transformer.proj_out``; attention.py:556-566, motion_module.py:72-75) so that MM-HAA and
the motion modules are numerically visible.
"""
import hashlib
import math

import torch


def _gen(seed: int, key: str) -> torch.Generator:
    h = hashlib.sha256(f"{seed}:{key}".encode()).digest()
    g = torch.Generator(device="cpu")
    g.manual_seed(int.from_bytes(h[:8], "little") & 0x7FFFFFFFFFFFFFFF)
    return g


def sinusoid_pe(max_len: int, d_model: int) -> torch.Tensor:
    """motion_module.py:262-277 PositionalEncoding buffer (1, max_len, d_model)."""
    pos = torch.arange(max_len).unsqueeze(1)
    div = torch.exp(torch.arange(0, d_model, 2) * (-math.log(10000.0) / d_model))
    pe = torch.zeros(1, max_len, d_model)
    pe[0, :, 0::2] = torch.sin(pos * div)
    pe[0, :, 1::2] = torch.cos(pos * div)
    return pe


def make_tensor(key: str, shape, seed: int = 0) -> torch.Tensor:
    shape = tuple(shape)
    g = _gen(seed, key)
    if key.endswith("pos_encoder.pe"):
        return sinusoid_pe(shape[1], shape[2])
    zero_init = ("zero_conv" in key) or ("temporal_transformer.proj_out" in key)
    if zero_init:
        return torch.randn(shape, generator=g) * 0.02
    is_norm = (".norm" in key or "conv_norm_out" in key or key.startswith("norm")
               or ".norms." in key or "ff_norm" in key)
    if is_norm and len(shape) == 1:
        if key.endswith("weight"):
            return 1.0 + 0.1 * torch.randn(shape, generator=g)
        return 0.1 * torch.randn(shape, generator=g)
    if len(shape) >= 2:
        fan_in = 1
        for s in shape[1:]:
            fan_in *= s
        bound = 1.0 / math.sqrt(fan_in)
        return (torch.rand(shape, generator=g) * 2 - 1) * bound
    # biases of linear / conv layers
    return (torch.rand(shape, generator=g) * 2 - 1) * 0.05


def make_state_dict(spec, seed: int = 0):
    """spec: iterable of (key, shape).  Returns {key: fp32 CPU tensor}."""
    return {k: make_tensor(k, s, seed) for k, s in spec}
