"""Host mirror of src/models/motion_module.py (AnimateDiff "Vanilla" temporal transformer).

Activations stay in the frame-major token layout (N = B*F, T, C) for the whole module: every Linear /
LayerNorm / FF is row-wise, and the temporal attention kernel gathers the F frames of each (batch, pixel)
itself, so none of the four ``(b f) d c <-> (b d) f c`` re-layouts of the reference
(motion_module.py:361-363,386) is materialised.
"""
import math

import torch
import torch.nn as nn

from .attention import Attention, FeedForward, _LN, ln_qkv
from .kernels import Engine
from .packing import Pack, f32, run
from .resnet import GroupNorm2d


def zero_module(module):
    for p in module.parameters():
        nn.init.zeros_(p)
    return module


def get_motion_module(in_channels, motion_module_type: str, motion_module_kwargs: dict):
    if motion_module_type == "Vanilla":
        return VanillaTemporalModule(in_channels=in_channels, **motion_module_kwargs)
    raise ValueError(motion_module_type)


class PositionalEncoding(nn.Module):
    """Sinusoidal table registered as the buffer ``pe`` (1, max_len, d_model) (motion_module.py:262-277)."""

    def __init__(self, d_model, dropout=0.0, max_len=24):
        super().__init__()
        position = torch.arange(max_len).unsqueeze(1)
        div_term = torch.exp(torch.arange(0, d_model, 2) * (-math.log(10000.0) / d_model))
        pe = torch.zeros(1, max_len, d_model)
        pe[0, :, 0::2] = torch.sin(position * div_term)
        pe[0, :, 1::2] = torch.cos(position * div_term)
        self.register_buffer("pe", pe)


class VersatileAttention(Attention):
    def __init__(self, attention_mode=None, cross_frame_attention_mode=None, temporal_position_encoding=False,
                 temporal_position_encoding_max_len=24, *args, **kwargs):
        super().__init__(*args, **kwargs)
        assert attention_mode == "Temporal"
        self.attention_mode = attention_mode
        self.is_cross_attention = kwargs.get("cross_attention_dim") is not None
        if self.is_cross_attention:
            raise NotImplementedError("Temporal_Cross blocks are not used by config/prompts/animation.yaml")
        self.pos_encoder = PositionalEncoding(kwargs["query_dim"], dropout=0.0, max_len=temporal_position_encoding_max_len) \
            if temporal_position_encoding else None
        self._pe_pack = Pack()

    def pe_table(self, eng: Engine):
        if self.pos_encoder is None:
            return None
        return self._pe_pack.get(eng, [self.pos_encoder.pe], lambda: f32(self.pos_encoder.pe[0], eng))


class TemporalTransformerBlock(nn.Module):
    def __init__(self, dim, num_attention_heads, attention_head_dim, attention_block_types=("Temporal_Self", "Temporal_Self"),
                 dropout=0.0, norm_num_groups=32, cross_attention_dim=768, activation_fn="geglu", attention_bias=False,
                 upcast_attention=False, cross_frame_attention_mode=None, temporal_position_encoding=False,
                 temporal_position_encoding_max_len=24):
        super().__init__()
        blocks, norms = [], []
        for name in attention_block_types:
            blocks.append(VersatileAttention(
                attention_mode=name.split("_")[0], cross_attention_dim=cross_attention_dim if name.endswith("_Cross") else None,
                query_dim=dim, heads=num_attention_heads, dim_head=attention_head_dim, dropout=dropout, bias=attention_bias,
                upcast_attention=upcast_attention, cross_frame_attention_mode=cross_frame_attention_mode,
                temporal_position_encoding=temporal_position_encoding,
                temporal_position_encoding_max_len=temporal_position_encoding_max_len))
            norms.append(_LN(dim))
        self.attention_blocks = nn.ModuleList(blocks)
        self.norms = nn.ModuleList(norms)
        self.ff = FeedForward(dim, dropout=dropout, activation_fn=activation_fn)
        self.ff_norm = _LN(dim)

    def run(self, eng: Engine, x, B: int, F: int, T: int, exchange=None):
        """x: (B*F*T, C) rows ordered (b, f, t) (motion_module.py:236-259, 351-388).  T may be a pixel chunk
        (token-sharded rows); ``exchange`` sends the block output back to the frame shards."""
        for attn, norm in zip(self.attention_blocks, self.norms):
            pk = attn.packed(eng)
            pe = attn.pos_encoder.pe[0] if attn.pos_encoder is not None else None
            if pe is not None and F > pe.shape[0]:
                raise ValueError(f"{F} frames per window exceed temporal_position_encoding_max_len={pe.shape[0]} "
                                 "(motion_module.py:275-277 fails the same way)")
            # LayerNorm, then + pe[frame]: the PE feeds q, k AND v (motion_module.py:365-366)
            qkv = ln_qkv(eng, x, norm, attn, pe=pe, T=T, F=F)
            a = eng.temporal_attention(qkv, B, F, T, attn.heads)
            x = eng.gemm(a, pk["o"], bias=pk["bo"], residual=x)
        return self.ff.run(eng, x, x, exchange=exchange, ln=self.ff_norm)


class TemporalTransformer3DModel(nn.Module):
    def __init__(self, in_channels, num_attention_heads, attention_head_dim, num_layers,
                 attention_block_types=("Temporal_Self", "Temporal_Self"), dropout=0.0, norm_num_groups=32,
                 cross_attention_dim=768, activation_fn="geglu", attention_bias=False, upcast_attention=False,
                 cross_frame_attention_mode=None, temporal_position_encoding=False, temporal_position_encoding_max_len=24):
        super().__init__()
        inner = num_attention_heads * attention_head_dim
        self.inner_dim = inner
        self.norm = GroupNorm2d(num_groups=norm_num_groups, num_channels=in_channels, eps=1e-6, affine=True)
        self.proj_in = nn.Linear(in_channels, inner)
        self.transformer_blocks = nn.ModuleList([
            TemporalTransformerBlock(dim=inner, num_attention_heads=num_attention_heads, attention_head_dim=attention_head_dim,
                                     attention_block_types=attention_block_types, dropout=dropout,
                                     norm_num_groups=norm_num_groups, cross_attention_dim=cross_attention_dim,
                                     activation_fn=activation_fn, attention_bias=attention_bias,
                                     upcast_attention=upcast_attention, cross_frame_attention_mode=cross_frame_attention_mode,
                                     temporal_position_encoding=temporal_position_encoding,
                                     temporal_position_encoding_max_len=temporal_position_encoding_max_len)
            for _ in range(num_layers)])
        self.proj_out = nn.Linear(inner, in_channels)
        self._pack = Pack()

    def run(self, eng: Engine, x, frames: int, shard=None):
        """x: (N, H, W, C) -> same (motion_module.py:146-182).  With ``shard`` (frame_shard.FrameShardGroup) x holds
        this rank's ``frames`` = F/k frames per sample: proj_in delivers its rows token-sharded to the k shards, the
        transformer blocks run on all F frames of T/k pixels, and the feed-forward's output projection delivers
        the rows back frame-sharded (one exchange each way per module)."""
        N, H, W, C = x.shape
        T = H * W
        B = N // frames
        if shard is not None and shard.k > 1:
            return self._run_sharded(eng, x, B, frames, T, shard)
        wi, bi, wo, bo = self._pack.get(
            eng, [self.proj_in.weight, self.proj_in.bias, self.proj_out.weight, self.proj_out.bias],
            lambda: (run(self.proj_in.weight, eng), f32(self.proj_in.bias, eng), run(self.proj_out.weight, eng),
                     f32(self.proj_out.bias, eng)))
        h = self.norm.run(eng, x, None, silu=False)
        tok = eng.gemm(h.view(N * T, C), wi, bias=bi)
        for blk in self.transformer_blocks:
            tok = blk.run(eng, tok, B, frames, T)
        out = eng.gemm(tok, wo, bias=bo, residual=x.view(N * T, C))
        return out.view(N, H, W, C)


    def _run_sharded(self, eng: Engine, x, B: int, frames_local: int, T: int, shard):
        N, H, W, C = x.shape
        k = shard.k
        F, Tc = frames_local * k, T // k
        wi, bi, wo, bo = self._pack.get(
            eng, [self.proj_in.weight, self.proj_in.bias, self.proj_out.weight, self.proj_out.bias],
            lambda: (run(self.proj_in.weight, eng), f32(self.proj_in.bias, eng), run(self.proj_out.weight, eng),
                     f32(self.proj_out.bias, eng)))
        h = self.norm.run(eng, x, None, silu=False)                      # per-frame GroupNorm: frame-local
        to_tokens = shard.exchange(1, B, F, T, self.inner_dim)
        eng.gemm(h.view(N * T, C), wi, bias=bi, exchange=to_tokens)      # rows land on the shard owning their pixel chunk
        shard.barrier()
        tok = to_tokens.recv                                             # (B*F*Tc, inner): rows (b, f, t_loc)
        to_frames = shard.exchange(2, B, F, T, self.inner_dim)
        last = len(self.transformer_blocks) - 1
        for i, blk in enumerate(self.transformer_blocks):
            tok = blk.run(eng, tok, B, F, Tc, exchange=to_frames if i == last else None)
        shard.barrier()
        out = eng.gemm(to_frames.recv, wo, bias=bo, residual=x.view(N * T, C))
        return out.view(N, H, W, C)


class VanillaTemporalModule(nn.Module):
    def __init__(self, in_channels, num_attention_heads=8, num_transformer_block=2,
                 attention_block_types=("Temporal_Self", "Temporal_Self"), cross_frame_attention_mode=None,
                 temporal_position_encoding=False, temporal_position_encoding_max_len=24, temporal_attention_dim_div=1,
                 zero_initialize=True):
        super().__init__()
        self.temporal_transformer = TemporalTransformer3DModel(
            in_channels=in_channels, num_attention_heads=num_attention_heads,
            attention_head_dim=in_channels // num_attention_heads // temporal_attention_dim_div,
            num_layers=num_transformer_block, attention_block_types=attention_block_types,
            cross_frame_attention_mode=cross_frame_attention_mode, temporal_position_encoding=temporal_position_encoding,
            temporal_position_encoding_max_len=temporal_position_encoding_max_len)
        if zero_initialize:
            self.temporal_transformer.proj_out = zero_module(self.temporal_transformer.proj_out)

    def run(self, eng: Engine, x, frames: int, shard=None):
        return self.temporal_transformer.run(eng, x, frames, shard)
