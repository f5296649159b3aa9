"""Host mirror of src/models/attention.py + the diffusers primitives it builds on.

``Attention`` / ``FeedForward`` here are parameter containers with the diffusers-0.24 state-dict layout
(``to_q/to_k/to_v.weight``, ``to_out.0.{weight,bias}``, ``net.0.proj``, ``net.2``); the math runs in the
CUDA kernels through ``run`` helpers on the block classes.
"""
from typing import List, Optional

import torch
import torch.nn as nn

from .kernels import Engine
from .packing import Pack, conv1x1, f32, geglu_interleave, ln_fold, run


class Attention(nn.Module):
    """diffusers Attention(query_dim, cross_attention_dim, heads, dim_head, bias=False, out_bias=True)."""

    def __init__(self, query_dim, cross_attention_dim=None, heads=8, dim_head=64, dropout=0.0, bias=False,
                 upcast_attention=False, **unused):
        super().__init__()
        if bias:
            raise NotImplementedError("attention_bias=True is not used by the reference config")
        inner = heads * dim_head
        self.heads, self.dim_head, self.inner_dim, self.query_dim = heads, dim_head, inner, query_dim
        self.kv_dim = cross_attention_dim if cross_attention_dim is not None else query_dim
        self.to_q = nn.Linear(query_dim, inner, bias=False)
        self.to_k = nn.Linear(self.kv_dim, inner, bias=False)
        self.to_v = nn.Linear(self.kv_dim, inner, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(inner, query_dim, bias=True), nn.Dropout(dropout)])
        self._pack = Pack()
        self._ln_pack = Pack()

    def packed_ln(self, eng: Engine, ln: nn.LayerNorm, pe: Optional[torch.Tensor] = None):
        """Self-attention q|k|v projection with the LayerNorm that feeds it folded in (packing.ln_fold):
        dict(w (3*inner, C) = [Wq;Wk;Wv] diag(gamma), colsum, bias = [Wq;Wk;Wv] beta, pe = pe [Wq;Wk;Wv]^T or None --
        the positional table of the motion module goes through the projection once, (max_len, 3*inner) float32, and is
        added per frame as a row bias (motion_module.py:365-366: the PE feeds q, k and v)."""
        assert self.kv_dim == self.query_dim
        srcs = [self.to_q.weight, self.to_k.weight, self.to_v.weight, ln.weight, ln.bias] + ([pe] if pe is not None else [])

        def build():
            w = torch.cat([self.to_q.weight, self.to_k.weight, self.to_v.weight], dim=0)
            wp, colsum, bias = ln_fold(w, None, ln.weight, ln.bias, eng)
            d = dict(w=wp, colsum=colsum, bias=bias, pe=None)
            if pe is not None:
                tab = pe.detach().to(device=eng.device, dtype=torch.float64) @ w.detach().to(device=eng.device, dtype=torch.float64).t()
                d["pe"] = tab.float().contiguous()
            return d
        return self._ln_pack.get(eng, srcs, build)

    def packed(self, eng: Engine):
        """-> dict(qkv (3*inner, C) if self-attention, q, kv (2*inner, kv_dim), o, bo)."""
        def build():
            d = {}
            q, k, v = self.to_q.weight, self.to_k.weight, self.to_v.weight
            if self.kv_dim == self.query_dim:
                d["qkv"] = run(torch.cat([q, k, v], dim=0), eng)
                d["q"] = d["qkv"][: self.inner_dim]
                d["kv"] = d["qkv"][self.inner_dim:]
            else:
                d["q"] = run(q, eng)
                d["kv"] = run(torch.cat([k, v], dim=0), eng)
            d["o"] = run(self.to_out[0].weight, eng)
            d["bo"] = f32(self.to_out[0].bias, eng)
            return d
        return self._pack.get(eng, [self.to_q.weight, self.to_k.weight, self.to_v.weight, self.to_out[0].weight,
                                    self.to_out[0].bias], build)


class GEGLU(nn.Module):
    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)


class FeedForward(nn.Module):
    """diffusers FeedForward(dim, activation_fn='geglu'): net = [GEGLU(dim, 4*dim), Dropout, Linear(4*dim, dim)]."""

    def __init__(self, dim, dropout=0.0, activation_fn="geglu", mult=4):
        super().__init__()
        if activation_fn != "geglu":
            raise NotImplementedError(activation_fn)
        self.net = nn.ModuleList([GEGLU(dim, dim * mult), nn.Dropout(dropout), nn.Linear(dim * mult, dim)])
        self._pack = Pack()
        self._ln_pack = Pack()

    def run(self, eng: Engine, x, residual, exchange=None, ln=None):
        """ff(LN(x)) + residual.  ``ln`` None: x (rows, dim) is already layer-normed.  With ``ln`` (the nn.LayerNorm in
        front of the feed-forward) the normalisation is folded into the GEGLU GEMM on the bf16 tensor-core tier (row
        statistics + epilogue), else it runs as its own pass first.  ``exchange`` (frame_shard.Exchange): the output
        rows are delivered to the shards that own them instead of being returned."""
        p1, p2 = self.net[0].proj, self.net[2]
        fold = ln is not None and eng.ln_fused
        if ln is not None and not fold:
            x = ln.run(eng, x)

        def build():
            gb = eng.geglu_block(p1.weight.shape[0])
            w1, b1 = geglu_interleave(p1.weight.detach(), p1.bias.detach(), gb)
            d = dict(gb=gb, w2=run(p2.weight, eng), b2=f32(p2.bias, eng))
            if fold:
                d["w1"], d["colsum"], d["b1"] = ln_fold(w1, b1, ln.weight, ln.bias, eng)
            else:
                d["w1"], d["b1"], d["colsum"] = run(w1, eng), f32(b1, eng), None
            return d
        srcs = [p1.weight, p1.bias, p2.weight, p2.bias] + ([ln.weight, ln.bias] if fold else [])
        pk = (self._ln_pack if fold else self._pack).get(eng, srcs, build)
        stats = eng.row_stats(x, ln.eps) if fold else None
        h = eng.gemm(x, pk["w1"], bias=pk["b1"], geglu_block=pk["gb"], rowstats=stats, colsum=pk["colsum"])
        return eng.gemm(h, pk["w2"], bias=pk["b2"], residual=residual, exchange=exchange)


class _LN(nn.LayerNorm):
    def __init__(self, dim):
        super().__init__(dim)
        self._pack = Pack()
        self._pe_pack = Pack()

    def run(self, eng: Engine, x, pe=None, T=0, F=0):
        g, b = self._pack.get(eng, [self.weight, self.bias], lambda: (f32(self.weight, eng), f32(self.bias, eng)))
        return eng.layernorm(x, g, b, self.eps, pe=pe, T=T, F=F)


def ln_qkv(eng: Engine, x, ln: "_LN", attn: Attention, pe=None, T: int = 0, F: int = 0):
    """LN(x) (+ pe[frame]) -> fused q|k|v projection (rows, 3*inner).  bf16 tensor-core tier: one pass over x for the row
    statistics, the normalisation itself happens in the GEMM epilogue; otherwise LayerNorm kernel + GEMM."""
    if eng.ln_fused:
        f = attn.packed_ln(eng, ln, pe)
        stats = eng.row_stats(x, ln.eps)
        if pe is None:
            return eng.gemm(x, f["w"], bias=f["bias"], rowstats=stats, colsum=f["colsum"])
        return eng.gemm(x, f["w"], bias=f["bias"], rowstats=stats, colsum=f["colsum"], rowbias=f["pe"][:F],
                        rows_per_group=T, rowbias_mod=F)
    pe_tab = None
    if pe is not None:
        pe_tab = ln._pe_pack.get(eng, [pe], lambda: f32(pe, eng))
    n = ln.run(eng, x, pe=pe_tab, T=T, F=F)
    return eng.gemm(n, attn.packed(eng)["qkv"])


class TemporalBasicTransformerBlock(nn.Module):
    """Spatial block of the denoising UNet: self-attention with ReferenceNet feature injection, CLIP
    cross-attention, GEGLU feed-forward.  The computation is the *read-mode hacked forward* that
    ReferenceAttentionControl installs (mutual_self_attention.py:93-230), not attention.py:382-481:

      x1 = attn1(LN1 x, kv = [LN1 x ; bank])  (+x)      frames flagged "uncond" use kv = LN1 x only -- the net
                                                         effect of the CFG re-do at :168-188, computed once
      x2 = attn2(LN2 x1, clip) + x1                      one CLIP token => softmax == 1 => to_out(to_v(clip))
      x3 = ff(LN3 x2) + x2

    ``bank`` is the list ReferenceAttentionControl.update() fills ((Bb, T, C) reference features)."""

    def __init__(self, dim, num_attention_heads, attention_head_dim, dropout=0.0, cross_attention_dim=None,
                 activation_fn="geglu", num_embeds_ada_norm=None, attention_bias=False, only_cross_attention=False,
                 upcast_attention=False, unet_use_cross_frame_attention=None, unet_use_temporal_attention=None,
                 name=None):
        super().__init__()
        if num_embeds_ada_norm is not None or only_cross_attention or unet_use_cross_frame_attention \
                or unet_use_temporal_attention:
            raise NotImplementedError("option not used by config/prompts/animation.yaml")
        self.name = name
        self.attn1 = Attention(dim, heads=num_attention_heads, dim_head=attention_head_dim, bias=attention_bias)
        self.norm1 = _LN(dim)
        self.attn2 = Attention(dim, cross_attention_dim=cross_attention_dim, heads=num_attention_heads,
                               dim_head=attention_head_dim, bias=attention_bias) if cross_attention_dim else None
        self.norm2 = _LN(dim) if cross_attention_dim else None
        self.ff = FeedForward(dim, dropout=dropout, activation_fn=activation_fn)
        self.norm3 = _LN(dim)
        self._bank_list: List[torch.Tensor] = []
        self._bank_gen = 0         # bumped whenever ``bank`` is assigned: a freed bank's address can be handed out again
        self._bank_kv = None       # (key, k2, v2, buffer) projected reference keys / values
        self.write_bank = False    # ReferenceNet write mode (mutual_self_attention.py:139-148): bank.append(norm1(x)), plain self-attention
        self._clip_pack = Pack()

    @property
    def bank(self) -> List[torch.Tensor]:
        return self._bank_list

    @bank.setter
    def bank(self, value):
        self._bank_list = value
        self._bank_gen += 1

    # -- reference K/V: the bank is constant over all steps and windows => project once
    def bank_kv(self, eng: Engine):
        if not self.bank:
            return None
        bank = self.bank[0]
        pk = self.attn1.packed(eng)
        key = (self._bank_gen, bank.data_ptr(), bank._version, tuple(bank.shape), pk["kv"].data_ptr(), str(eng.device),
               eng.dtype)
        if self._bank_kv is None or self._bank_kv[0] != key:
            Bb, T, C = bank.shape
            b = bank.to(device=eng.device, dtype=eng.dtype).contiguous().view(Bb * T, C)
            old = self._bank_kv
            if old is not None and tuple(old[3].shape) == (Bb, T, 2 * C) and old[3].dtype == eng.dtype \
                    and old[3].device == eng.device:
                # a new reference image of the same shape: project IN PLACE -- a CUDA graph captured for the previous
                # video reads this storage (DenoiseLoop.reload)
                kv = old[3]
                eng.gemm(b, pk["kv"], out=kv.view(Bb * T, 2 * C))
            else:
                kv = eng.gemm(b, pk["kv"]).view(Bb, T, 2 * C)
            self._bank_kv = (key, kv[:, :, :C], kv[:, :, C:], kv)
        return self._bank_kv[1], self._bank_kv[2]

    def bank_kv_storage(self):
        """data_ptr of the projected reference K/V buffer (None before the first projection)."""
        return None if self._bank_kv is None else self._bank_kv[3].data_ptr()

    def clip_vector(self, eng: Engine, clip_b):
        """attn2 with a single key/value token: (B, 1, 768) -> (B, C) float32 = to_out(to_v(clip))."""
        a = self.attn2
        wv, wo, bo = self._clip_pack.get(
            eng, [a.to_v.weight, a.to_out[0].weight, a.to_out[0].bias],
            lambda: (f32(a.to_v.weight, eng), f32(a.to_out[0].weight, eng), f32(a.to_out[0].bias, eng)))
        v = eng.gemm(clip_b.reshape(clip_b.shape[0], -1).float().contiguous(), wv, dtype=torch.float32)
        return eng.gemm(v, wo, bias=bo, dtype=torch.float32)

    def run(self, eng: Engine, tok, clip_b, frames: int, seg2_index):
        """tok: (N, T, C); clip_b: (B, L_clip, 768); seg2_index: (N,) int32 bank row per frame or -1."""
        N, T, C = tok.shape
        rows = N * T
        heads = self.attn1.heads
        x = tok.view(rows, C)
        pk = self.attn1.packed(eng)
        if self.write_bank:
            # write mode: the normalised hidden states ARE the reference features (one (B, T, C) tensor per block)
            n1 = self.norm1.run(eng, x)
            self._bank_list.append(n1.view(N, T, C).clone())
            qkv = eng.gemm(n1, pk["qkv"]).view(N, T, 3 * C)
            bkv = None
        else:
            qkv = ln_qkv(eng, x, self.norm1, self.attn1).view(N, T, 3 * C)
            bkv = self.bank_kv(eng)
        k2, v2 = bkv if bkv is not None else (None, None)
        a = eng.attention(qkv[:, :, :C], qkv[:, :, C:2 * C], qkv[:, :, 2 * C:], heads, k2=k2, v2=v2,
                          seg2_index=seg2_index if bkv is not None else None)
        if self.attn2 is not None and clip_b.shape[1] == 1:
            cvec = self.clip_vector(eng, clip_b)
            x = eng.gemm(a.view(rows, C), pk["o"], bias=pk["bo"], residual=x, rowbias=cvec, rows_per_group=frames * T)
        else:
            x = eng.gemm(a.view(rows, C), pk["o"], bias=pk["bo"], residual=x)
            if self.attn2 is not None:
                x = self._cross_general(eng, x, clip_b, N, T, frames)
        return self.ff.run(eng, x, x, ln=self.norm3).view(N, T, C)

    def _cross_general(self, eng, x, ctx_b, N, T, frames):
        """attn2 for more than one context token (not exercised by the pipeline, kept for API parity)."""
        a = self.attn2
        pk = a.packed(eng)
        C = x.shape[-1]
        B, Lc, Dc = ctx_b.shape
        n2 = self.norm2.run(eng, x)
        q = eng.gemm(n2, pk["q"]).view(N, T, C)
        kv = eng.gemm(ctx_b.to(eng.dtype).contiguous().view(B * Lc, Dc), pk["kv"]).view(B, Lc, 2 * C)
        idx = torch.arange(B, device=x.device).repeat_interleave(frames)
        kvn = kv[idx].contiguous()
        o = eng.attention(q, kvn[:, :, :C], kvn[:, :, C:], a.heads)
        return eng.gemm(o.view(N * T, C), pk["o"], bias=pk["bo"], residual=x)


class AudioTemporalBasicTransformerBlock(nn.Module):
    """MM-HAA: spatial self-attention, three audio cross-attentions gated by the full / face / lip motion
    masks, zero-initialised 1x1 convs, weighted hierarchical sum, GEGLU FF (attention.py:486-771)."""

    def __init__(self, dim, num_attention_heads, attention_head_dim, dropout=0.0, cross_attention_dim=None,
                 activation_fn="geglu", num_embeds_ada_norm=None, attention_bias=False, only_cross_attention=False,
                 upcast_attention=False, unet_use_cross_frame_attention=None, unet_use_temporal_attention=None, depth=0,
                 unet_block_name=None, stack_enable_blocks_name=None, stack_enable_blocks_depth=None):
        super().__init__()
        if num_embeds_ada_norm is not None or unet_use_cross_frame_attention:
            raise NotImplementedError("option not used by config/prompts/animation.yaml")
        if not (cross_attention_dim is not None and stack_enable_blocks_name is not None
                and stack_enable_blocks_depth is not None and unet_block_name in stack_enable_blocks_name
                and depth in stack_enable_blocks_depth):
            raise NotImplementedError("AudioTemporalBasicTransformerBlock without the 3-branch MM-HAA stack")
        self.depth = depth
        self.unet_block_name = unet_block_name
        self.zero_conv_full = nn.Conv2d(dim, dim, kernel_size=1)
        self.zero_conv_face = nn.Conv2d(dim, dim, kernel_size=1)
        self.zero_conv_lip = nn.Conv2d(dim, dim, kernel_size=1)
        for m in (self.zero_conv_full, self.zero_conv_face, self.zero_conv_lip):   # zero_module (attention.py:773-785)
            nn.init.zeros_(m.weight)
            nn.init.zeros_(m.bias)
        mk = lambda ca: Attention(dim, cross_attention_dim=ca, heads=num_attention_heads,  # noqa: E731
                                  dim_head=attention_head_dim, bias=attention_bias)
        self.attn1 = mk(None)
        self.norm1 = _LN(dim)
        self.attn2_0, self.attn2_1, self.attn2_2 = mk(cross_attention_dim), mk(cross_attention_dim), mk(cross_attention_dim)
        self.attn2 = None
        self.norm2 = _LN(dim)
        self.ff = FeedForward(dim, dropout=dropout, activation_fn=activation_fn)
        self.norm3 = _LN(dim)
        self._pack = Pack()
        self._q3_pack = Pack()
        self.fuse_regions = True       # bf16 tier: fused three-region kernel + one GEMM (False = per-region operators)
        self._fused_key, self._fused_val = None, None

    def _packed(self, eng: Engine):
        branches = (self.attn2_0, self.attn2_1, self.attn2_2)
        zcs = (self.zero_conv_full, self.zero_conv_face, self.zero_conv_lip)
        params = [p for a in branches for p in (a.to_q.weight, a.to_k.weight, a.to_v.weight, a.to_out[0].weight,
                                                 a.to_out[0].bias)] + [p for z in zcs for p in (z.weight, z.bias)]

        def build():
            d = {}
            d["q3"] = run(torch.cat([a.to_q.weight for a in branches], dim=0), eng)                    # (3C, C)
            d["kv6"] = run(torch.cat([w for a in branches for w in (a.to_k.weight, a.to_v.weight)], dim=0), eng)
            d["o"] = [run(a.to_out[0].weight, eng) for a in branches]
            d["bo"] = [f32(a.to_out[0].bias, eng) for a in branches]
            d["z"] = [conv1x1(z.weight, eng) for z in zcs]
            d["bz"] = [f32(z.bias, eng) for z in zcs]
            return d
        return self._pack.get(eng, params, build)

    def _q3_ln(self, eng: Engine):
        """The three audio query projections [Wq_full; Wq_face; Wq_lip] with norm2 folded in (packing.ln_fold)."""
        branches = (self.attn2_0, self.attn2_1, self.attn2_2)
        srcs = [a.to_q.weight for a in branches] + [self.norm2.weight, self.norm2.bias]

        def build():
            w, colsum, bias = ln_fold(torch.cat([a.to_q.weight for a in branches], dim=0), None, self.norm2.weight,
                                      self.norm2.bias, eng)
            return dict(w=w, colsum=colsum, bias=bias)
        return self._q3_pack.get(eng, srcs, build)

    def _fused_regions(self, eng: Engine, pk, scale):
        """W' = [Wz_0 Wo_0 | Wz_1 Wo_1 | Wz_2 Wo_2 | Wz_0 bo_0, Wz_1 bo_1, Wz_2 bo_2, 0 x 5]  (C, 3C + 8) and
        bias' = sum_r scale_r bz_r: with A' = [gate_r * attn_r | gate_r | 0] (mmgt_audio_attention),
        A' W'^T + bias' = sum_r scale_r * zero_conv_r(mask_r * to_out_r(attn_r)), attention.py:719-767."""
        key = (id(pk), scale)
        if self._fused_key != key:
            branches = (self.attn2_0, self.attn2_1, self.attn2_2)
            zcs = (self.zero_conv_full, self.zero_conv_face, self.zero_conv_lip)
            with torch.no_grad():
                cols, tails, bias = [], [], 0.0
                for a, z, s in zip(branches, zcs, scale):
                    wz = z.weight.detach().double().reshape(z.weight.shape[0], -1)
                    cols.append(wz @ a.to_out[0].weight.detach().double())
                    tails.append(wz @ a.to_out[0].bias.detach().double())
                    bias = bias + s * z.bias.detach().double()
                Cc = cols[0].shape[0]
                w = torch.cat(cols + [torch.stack(tails, dim=1), torch.zeros(Cc, 5, dtype=torch.float64, device=cols[0].device)],
                              dim=1)
                self._fused_val = dict(w=run(w.float(), eng), bias=f32(bias.float(), eng))
            self._fused_key = key
        return self._fused_val

    def run(self, eng: Engine, tok, audio_rows, masks, scale):
        """tok (N,T,C); audio_rows (N*M, 768) run dtype; masks: 3 x (N*T,) float32; scale: 3 floats."""
        N, T, C = tok.shape
        rows = N * T
        heads = self.attn1.heads
        x = tok.view(rows, C)
        p1 = self.attn1.packed(eng)
        qkv = ln_qkv(eng, x, self.norm1, self.attn1).view(N, T, 3 * C)
        a = eng.attention(qkv[:, :, :C], qkv[:, :, C:2 * C], qkv[:, :, 2 * C:], heads)
        x = eng.gemm(a.view(rows, C), p1["o"], bias=p1["bo"], residual=x)
        pk = self._packed(eng)
        if eng.ln_fused:
            f = self._q3_ln(eng)
            q3 = eng.gemm(x, f["w"], bias=f["bias"], rowstats=eng.row_stats(x, self.norm2.eps), colsum=f["colsum"]).view(N, T, 3 * C)
        else:
            n2 = self.norm2.run(eng, x)
            q3 = eng.gemm(n2, pk["q3"]).view(N, T, 3 * C)
        M = audio_rows.shape[0] // N
        if self.fuse_regions and eng.audio_attention_supported(M, C // heads):
            # One kernel for the three audio cross-attentions with the mask gate and motion_scale in its epilogue, then
            # ONE GEMM (K = 3C + 8) for to_out_r -> zero_conv_r -> weighted sum -> + x (weights pre-multiplied).
            fz = self._fused_regions(eng, pk, tuple(float(s) for s in scale))
            kv6 = eng.gemm(audio_rows, pk["kv6"])
            gated = eng.audio_attention(q3.view(rows, 3 * C), kv6, masks, scale, N, T, heads)
            x = eng.gemm(gated, fz["w"], bias=fz["bias"], residual=x)
            return self.ff.run(eng, x, x, ln=self.norm3).view(N, T, C)
        kv6 = eng.gemm(audio_rows, pk["kv6"]).view(N, M, 6 * C)
        acc = x
        for r in range(3):
            o = eng.attention(q3[:, :, r * C:(r + 1) * C], kv6[:, :, 2 * r * C:(2 * r + 1) * C],
                              kv6[:, :, (2 * r + 1) * C:(2 * r + 2) * C], heads)
            y = eng.gemm(o.view(rows, C), pk["o"][r], bias=pk["bo"][r], rowscale=masks[r])
            acc = eng.gemm(y, pk["z"][r], bias=pk["bz"][r], alpha=float(scale[r]), residual=acc)
        x = acc
        return self.ff.run(eng, x, x, ln=self.norm3).view(N, T, C)


def zero_module(module):
    for p in module.parameters():
        nn.init.zeros_(p)
    return module
