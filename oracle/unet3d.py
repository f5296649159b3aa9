"""CPU restatement of one ``UNet3DConditionModel`` step (MMGT stage 2).  TEST INFRASTRUCTURE.

Functional, state-dict driven, plain PyTorch fp32.  It is written against the reference's
*behaviour* (file:line cited per function, paths relative to /root/reference) and is pinned to
the reference's own modules by ``oracle/make_golden.py`` + ``tests/test_oracle_golden.py``.

Layout conventions used here (not the reference's): frames are always flattened ``n = b*F + f``;
conv-side tensors are ``(N, C, H, W)``, token-side tensors ``(N, T, C)`` with ``T = H*W``.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F


@dataclass
class UNetSpec:
    """The subset of the reference config the hot path depends on (unet_3d.py:37-90)."""
    block_out_channels: Sequence[int] = (320, 640, 1280, 1280)
    layers_per_block: int = 2
    heads: int = 8                     # `attention_head_dim=8` is used as head COUNT (unet_3d.py:150)
    norm_num_groups: int = 32
    norm_eps: float = 1e-5             # resnets + conv_norm_out; transformer/motion norms use 1e-6
    cross_attention_dim: int = 768
    audio_attention_dim: int = 768
    in_channels: int = 4
    out_channels: int = 4
    flip_sin_to_cos: bool = True
    freq_shift: int = 0
    pe_max_len: int = 32
    audio_levels: Sequence[int] = (0, 1, 2)   # MM-HAA only in the 3 CrossAttnDownBlock3D (unet_3d.py:165-169)


class _SD:
    """Prefix view over a flat state dict."""

    def __init__(self, sd: Dict[str, torch.Tensor], prefix: str = ""):
        self.sd, self.prefix = sd, prefix

    def sub(self, name: str) -> "_SD":
        return _SD(self.sd, f"{self.prefix}{name}.")

    def __getitem__(self, name: str) -> torch.Tensor:
        return self.sd[self.prefix + name]

    def has(self, name: str) -> bool:
        return (self.prefix + name) in self.sd


# --------------------------------------------------------------------------- primitives

def linear(p: _SD, x):
    return F.linear(x, p["weight"], p["bias"] if p.has("bias") else None)


def conv2d(p: _SD, x, stride=1, padding=1):
    return F.conv2d(x, p["weight"], p["bias"], stride=stride, padding=padding)


def group_norm(p: _SD, x, groups, eps):
    return F.group_norm(x, groups, p["weight"], p["bias"], eps)


def layer_norm(p: _SD, x):
    return F.layer_norm(x, (x.shape[-1],), p["weight"], p["bias"], 1e-5)


def sdpa(q, k, v, heads):
    """diffusers 0.24 AttnProcessor2_0 core: softmax(q k^T / sqrt(d)) v per head (App. A)."""
    n, lq, c = q.shape
    d = c // heads
    qh = q.view(n, lq, heads, d).transpose(1, 2)
    kh = k.view(n, -1, heads, d).transpose(1, 2)
    vh = v.view(n, -1, heads, d).transpose(1, 2)
    o = F.scaled_dot_product_attention(qh, kh, vh)   # softmax(q k^T / sqrt(d)) v, what AttnProcessor2_0 calls
    return o.transpose(1, 2).reshape(n, lq, c)


def attention(p: _SD, x, kv, heads):
    """diffusers Attention.forward: to_q / to_k / to_v (no bias) -> SDPA -> to_out.0 (bias)."""
    q = F.linear(x, p["to_q.weight"])
    k = F.linear(kv, p["to_k.weight"])
    v = F.linear(kv, p["to_v.weight"])
    return linear(p.sub("to_out.0"), sdpa(q, k, v, heads))


def feed_forward(p: _SD, x):
    """diffusers FeedForward(geglu): value half first, gate half second, exact erf GELU."""
    h = linear(p.sub("net.0.proj"), x)
    val, gate = h.chunk(2, dim=-1)
    return linear(p.sub("net.2"), val * F.gelu(gate))


def timestep_embedding(spec: UNetSpec, sd: _SD, timestep, batch):
    """diffusers Timesteps(320, flip=True, shift=0) + TimestepEmbedding (unet_3d.py:481-502)."""
    t = torch.as_tensor(timestep).reshape(-1).to(torch.float32)
    if t.numel() == 1:
        t = t.expand(batch)
    dim = spec.block_out_channels[0]
    half = dim // 2
    freq = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32) / (half - spec.freq_shift))
    arg = t[:, None] * freq[None, :]
    emb = torch.cat([torch.cos(arg), torch.sin(arg)] if spec.flip_sin_to_cos
                    else [torch.sin(arg), torch.cos(arg)], dim=-1)
    emb = linear(sd.sub("time_embedding.linear_1"), emb)
    return linear(sd.sub("time_embedding.linear_2"), F.silu(emb))


# --------------------------------------------------------------------------- blocks

def resnet_block(spec: UNetSpec, p: _SD, x, emb_n):
    """ResnetBlock3D.forward (resnet.py:217-247); emb_n is the per-frame (N, 1280) embedding."""
    h = F.silu(group_norm(p.sub("norm1"), x, spec.norm_num_groups, spec.norm_eps))
    h = conv2d(p.sub("conv1"), h)
    h = h + linear(p.sub("time_emb_proj"), F.silu(emb_n))[:, :, None, None]
    h = F.silu(group_norm(p.sub("norm2"), h, spec.norm_num_groups, spec.norm_eps))
    h = conv2d(p.sub("conv2"), h)
    if p.has("conv_shortcut.weight"):
        x = conv2d(p.sub("conv_shortcut"), x, padding=0)
    return x + h  # output_scale_factor == 1 (unet_3d.py:181)


def _tokens_in(p: _SD, x, groups):
    """Transformer3DModel entry: GroupNorm(eps 1e-6) + 1x1 conv, to tokens (transformer_3d.py:171-180)."""
    h = group_norm(p.sub("norm"), x, groups, 1e-6)
    h = conv2d(p.sub("proj_in"), h, padding=0)
    n, c, hh, ww = h.shape
    return h.permute(0, 2, 3, 1).reshape(n, hh * ww, c)


def _tokens_out(p: _SD, tok, x):
    """Transformer3DModel exit: tokens -> 1x1 conv proj_out -> + residual (transformer_3d.py:247-262)."""
    n, _, hh, ww = x.shape
    h = tok.reshape(n, hh, ww, tok.shape[-1]).permute(0, 3, 1, 2)
    return conv2d(p.sub("proj_out"), h, padding=0) + x


def spatial_transformer(spec: UNetSpec, p: _SD, x, clip_n, bank_n: Optional[torch.Tensor],
                        use_ref: torch.Tensor):
    """Transformer3DModel + the *hacked* TemporalBasicTransformerBlock in read mode
    (transformer_3d.py:139-268, mutual_self_attention.py:93-230).

    bank_n : (N, T, C) reference features already expanded per frame (or None)
    use_ref: (N,) bool; frames with False get plain self-attention -- the net effect of the CFG
             "uncond re-do" at mutual_self_attention.py:168-188.
    clip_n : (N, 1, 768) CLIP image embedding per frame.
    """
    tok = _tokens_in(p, x, spec.norm_num_groups)
    b = p.sub("transformer_blocks.0")
    n1 = layer_norm(b.sub("norm1"), tok)
    if bank_n is not None and bool(use_ref.any()):
        a1 = torch.empty_like(n1)
        ref_rows, self_rows = use_ref.nonzero().flatten(), (~use_ref).nonzero().flatten()
        a1[ref_rows] = attention(b.sub("attn1"), n1[ref_rows], torch.cat([n1[ref_rows], bank_n[ref_rows]], dim=1), spec.heads)
        if self_rows.numel():
            a1[self_rows] = attention(b.sub("attn1"), n1[self_rows], n1[self_rows], spec.heads)
    else:
        a1 = attention(b.sub("attn1"), n1, n1, spec.heads)
    tok = a1 + tok
    tok = attention(b.sub("attn2"), layer_norm(b.sub("norm2"), tok), clip_n, spec.heads) + tok
    tok = feed_forward(b.sub("ff"), layer_norm(b.sub("norm3"), tok)) + tok
    return _tokens_out(p, tok, x)


def audio_transformer(spec: UNetSpec, p: _SD, x, audio_n, masks, scale):
    """Transformer3DModel(use_audio_module) + AudioTemporalBasicTransformerBlock = MM-HAA
    (transformer_3d.py:160-164,234-244; attention.py:649-771).

    audio_n: (N, 32, 768); masks = (full, face, lip) each (N, T) for this block's level;
    scale = motion_scale (3 floats) or (1,1,1) when the eval branch drops it (unet_3d_blocks.py:591-600).
    """
    tok = _tokens_in(p, x, spec.norm_num_groups)
    b = p.sub("transformer_blocks.0")
    n1 = layer_norm(b.sub("norm1"), tok)
    tok = attention(b.sub("attn1"), n1, n1, spec.heads) + tok
    n2 = layer_norm(b.sub("norm2"), tok)
    acc = tok
    for r, (attn_name, zc_name) in enumerate((("attn2_0", "zero_conv_full"), ("attn2_1", "zero_conv_face"),
                                              ("attn2_2", "zero_conv_lip"))):
        h = attention(b.sub(attn_name), n2, audio_n, spec.heads) * masks[r][:, :, None]
        zc = b.sub(zc_name)
        h = F.linear(h, zc["weight"][:, :, 0, 0], zc["bias"])  # 1x1 conv on the token grid
        acc = acc + scale[r] * h
    # NB reference sums (s0*full + s1*face + s2*lip + x); same value up to fp32 association
    tok = acc
    tok = feed_forward(b.sub("ff"), layer_norm(b.sub("norm3"), tok)) + tok
    return _tokens_out(p, tok, x)


def motion_module(spec: UNetSpec, p: _SD, x, frames: int):
    """VanillaTemporalModule (motion_module.py:77-91,146-182,236-259,351-388)."""
    t = p.sub("temporal_transformer")
    n, c, hh, ww = x.shape
    bsz = n // frames
    h = group_norm(t.sub("norm"), x, spec.norm_num_groups, 1e-6)
    tok = h.permute(0, 2, 3, 1).reshape(n, hh * ww, c)
    tok = linear(t.sub("proj_in"), tok)
    blk = t.sub("transformer_blocks.0")
    for i in range(2):
        nrm = layer_norm(blk.sub(f"norms.{i}"), tok)
        ab = blk.sub(f"attention_blocks.{i}")
        # (b f) d c -> (b d) f c ; PE added to the normed states feeding q, k AND v (motion_module.py:365-366)
        seq = nrm.reshape(bsz, frames, hh * ww, c).permute(0, 2, 1, 3).reshape(bsz * hh * ww, frames, c)
        seq = seq + ab["pos_encoder.pe"][:, :frames]
        o = attention(ab, seq, seq, spec.heads)
        o = o.reshape(bsz, hh * ww, frames, c).permute(0, 2, 1, 3).reshape(n, hh * ww, c)
        tok = o + tok
    tok = feed_forward(blk.sub("ff"), layer_norm(blk.sub("ff_norm"), tok)) + tok
    tok = linear(t.sub("proj_out"), tok)
    return tok.reshape(n, hh, ww, c).permute(0, 3, 1, 2) + x


# --------------------------------------------------------------------------- whole model

def spatial_block_prefixes(spec: UNetSpec) -> List[str]:
    """Prefixes of the 16 spatial transformer blocks in the reference's module (DFS) order.

    NB the order is down -> up -> mid: ``self.mid_block = None`` (unet_3d.py:118) is a plain attribute
    until the real module is assigned at :176, i.e. after ``up_blocks`` was registered at :119."""
    out = []
    for i in range(3):
        for j in range(spec.layers_per_block):
            out.append(f"down_blocks.{i}.attentions.{j}")
    for i in range(1, 4):
        for j in range(spec.layers_per_block + 1):
            out.append(f"up_blocks.{i}.attentions.{j}")
    out.append("mid_block.attentions.0")
    return out


def spatial_block_width(spec: UNetSpec, prefix: str) -> int:
    boc = list(spec.block_out_channels)
    parts = prefix.split(".")
    if parts[0] == "down_blocks":
        return boc[int(parts[1])]
    if parts[0] == "mid_block":
        return boc[-1]
    return list(reversed(boc))[int(parts[1])]


def bank_pairing_order(spec: UNetSpec) -> List[str]:
    """Order in which ReferenceAttentionControl pairs reader/writer blocks: stable sort of the DFS
    order by descending norm1 width (mutual_self_attention.py:286-288,333-340)."""
    pref = spatial_block_prefixes(spec)
    return sorted(pref, key=lambda s: -spatial_block_width(spec, s))


def unet3d_forward(sd: Dict[str, torch.Tensor], spec: UNetSpec, sample, timestep, encoder_hidden_states,
                   audio_embedding, pose_cond_fea, full_mask, face_mask, body_mask, motion_scale,
                   banks: Dict[str, torch.Tensor], ref_index: Sequence[Optional[int]],
                   apply_motion_scale: bool = True, taps: Optional[dict] = None):
    """UNet3DConditionModel.forward (unet_3d.py:425-625) for one context window.

    sample (B,4,F,H,W); encoder_hidden_states (B,1,768); audio_embedding (B,F,32,768);
    pose_cond_fea (B,320,F,H,W); *_mask: list over levels of (B*F, T_level);
    banks: {spatial block prefix: (Bb, T_level, C)}; ref_index[b] = bank row used by sample b, or None
    for "self-attention only" (the CFG uncond half).  ``taps`` (optional dict) receives the output of
    every sub-module keyed by its state-dict prefix, in NCFHW, for golden localisation.
    """
    root = _SD(sd)
    B, _, Fr, H, W = sample.shape
    N = B * Fr
    boc = list(spec.block_out_channels)

    def to_frames(t5):  # (B,C,F,H,W) -> (N,C,H,W)
        return t5.permute(0, 2, 1, 3, 4).reshape(N, t5.shape[1], t5.shape[3], t5.shape[4])

    def tap(name, x):
        if taps is not None:
            taps[name] = x.reshape(B, Fr, *x.shape[1:]).permute(0, 2, 1, 3, 4).clone()

    emb = timestep_embedding(spec, root, timestep, B)                 # (B, 1280)
    emb_n = emb.repeat_interleave(Fr, dim=0)                           # (N, 1280)
    clip_n = encoder_hidden_states.repeat_interleave(Fr, dim=0)        # (N, 1, 768)
    audio_n = audio_embedding.reshape(N, audio_embedding.shape[2], audio_embedding.shape[3])
    use_ref = torch.tensor([ref_index[b] is not None for b in range(B)]).repeat_interleave(Fr)
    scale = list(motion_scale) if (motion_scale is not None and apply_motion_scale) else [1.0, 1.0, 1.0]

    def bank_for(prefix):
        if banks is None or prefix not in banks:
            return None
        bk = banks[prefix]
        rows = [bk[ref_index[b] if ref_index[b] is not None else 0] for b in range(B)]
        return torch.stack(rows).repeat_interleave(Fr, dim=0)          # (N, T, C)

    x = conv2d(root.sub("conv_in"), to_frames(sample))
    if pose_cond_fea is not None:
        x = x + to_frames(pose_cond_fea)
    tap("conv_in", x)
    skips = [x]

    # ---- down (unet_3d_blocks.py:519-618, 694-737)
    for i in range(4):
        blk = root.sub(f"down_blocks.{i}")
        for j in range(spec.layers_per_block):
            x = resnet_block(spec, blk.sub(f"resnets.{j}"), x, emb_n)
            tap(f"down_blocks.{i}.resnets.{j}", x)
            if i < 3:
                pre = f"down_blocks.{i}.attentions.{j}"
                x = spatial_transformer(spec, root.sub(pre), x, clip_n, bank_for(pre), use_ref)
                tap(pre, x)
                if i in spec.audio_levels:
                    x = audio_transformer(spec, blk.sub(f"audio_modules.{j}"), x, audio_n,
                                          (full_mask[i], face_mask[i], body_mask[i]), scale)
                    tap(f"down_blocks.{i}.audio_modules.{j}", x)
            x = motion_module(spec, blk.sub(f"motion_modules.{j}"), x, Fr)
            tap(f"down_blocks.{i}.motion_modules.{j}", x)
            skips.append(x)
        if i < 3:
            x = conv2d(blk.sub("downsamplers.0.conv"), x, stride=2, padding=1)
            tap(f"down_blocks.{i}.downsamplers.0", x)
            skips.append(x)

    # ---- mid (unet_3d_blocks.py:330-378); no audio module here (fact 2)
    mid = root.sub("mid_block")
    x = resnet_block(spec, mid.sub("resnets.0"), x, emb_n)
    pre = "mid_block.attentions.0"
    x = spatial_transformer(spec, root.sub(pre), x, clip_n, bank_for(pre), use_ref)
    tap(pre, x)
    x = motion_module(spec, mid.sub("motion_modules.0"), x, Fr)
    x = resnet_block(spec, mid.sub("resnets.1"), x, emb_n)
    tap("mid_block", x)

    # ---- up (unet_3d_blocks.py:872-975, 1045-1092)
    for i in range(4):
        blk = root.sub(f"up_blocks.{i}")
        for j in range(spec.layers_per_block + 1):
            x = torch.cat([x, skips.pop()], dim=1)
            x = resnet_block(spec, blk.sub(f"resnets.{j}"), x, emb_n)
            tap(f"up_blocks.{i}.resnets.{j}", x)
            if i > 0:
                pre = f"up_blocks.{i}.attentions.{j}"
                x = spatial_transformer(spec, root.sub(pre), x, clip_n, bank_for(pre), use_ref)
                tap(pre, x)
            x = motion_module(spec, blk.sub(f"motion_modules.{j}"), x, Fr)
            tap(f"up_blocks.{i}.motion_modules.{j}", x)
        if i < 3:
            x = F.interpolate(x, scale_factor=2.0, mode="nearest")    # resnet.py:70-73
            x = conv2d(blk.sub("upsamplers.0.conv"), x)
            tap(f"up_blocks.{i}.upsamplers.0", x)

    # ---- out (unet_3d.py:618-620)
    x = F.silu(group_norm(root.sub("conv_norm_out"), x, spec.norm_num_groups, spec.norm_eps))
    x = conv2d(root.sub("conv_out"), x)
    return x.reshape(B, Fr, x.shape[1], H, W).permute(0, 2, 1, 3, 4).contiguous()
