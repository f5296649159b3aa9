"""A CPU stand-in for mmgt_b200.kernels.Engine, for the `-m "not gpu"` tests ONLY.

It implements the operator contracts of include/mmgt_b200.h in plain float32 PyTorch so that the HOST side of the product
(module orchestration, weight packs, GEGLU interleave, fused MM-HAA weights, window gather / accumulate, CFG + DDIM, frame-shard
row exchange) can be executed and compared with the oracle where there is no GPU.  It is test infrastructure like oracle/:
nothing under mmgt_b200/ imports it, and it says nothing about the CUDA kernels (those are covered by the -m gpu tests).
"""
import math
from typing import List, Optional

import torch
import torch.nn.functional as F


class _Ctx:
    def tensor_cores(self):
        return True

    def launches(self):
        return 0


class _Lib:
    @staticmethod
    def mmgt_gemm_tc_block_n(n):
        for c in (256, 160, 128, 64, 32):
            if n % c == 0:
                return c
        return 0


class FakeEngine:
    def __init__(self, fuse_audio: bool = True, interleaved_geglu: bool = True, ln_fused: bool = False):
        self.device = torch.device("cpu")
        self.dtype = torch.float32
        self.subpixel_upsample = ln_fused   # and the sub-pixel form of Upsample3D (packing.subpixel_pack)
        self.ln_fused = ln_fused        # exercise the host's LayerNorm folding (packing.ln_fold) against plain LayerNorm
        self.ctx, self.lib, self.h, self.prof = _Ctx(), _Lib(), None, None
        self.unfused_exchange = False
        self.fuse_audio, self.interleaved_geglu = fuse_audio, interleaved_geglu
        self.calls = {}

    def _count(self, name):
        self.calls[name] = self.calls.get(name, 0) + 1

    # ------------------------------------------------------------------ helpers
    def empty(self, *shape, dtype=None):
        return torch.empty(shape, dtype=dtype or self.dtype)

    def geglu_block(self, n_rows: int) -> int:
        return 16 if (self.interleaved_geglu and n_rows % 32 == 0) else n_rows // 2

    # ------------------------------------------------------------------ layout
    def ncfhw_to_tokens(self, x, add=None):
        B, C, Fr, H, W = x.shape
        x = x.float() + (add.float() if add is not None else 0.0)
        return x.permute(0, 2, 3, 4, 1).reshape(B * Fr, H, W, C).contiguous()

    def tokens_to_ncfhw(self, x, B, Fr, out_dtype):
        N, H, W, C = x.shape
        return x.reshape(B, Fr, H, W, C).permute(0, 4, 1, 2, 3).contiguous().to(out_dtype)

    # ------------------------------------------------------------------ norms
    def groupnorm(self, x1, x2, gamma, beta, groups, eps, silu):
        self._count("groupnorm")
        x = x1 if x2 is None else torch.cat([x1, x2], dim=-1)
        shp = x.shape
        N, C = shp[0], shp[-1]
        y = F.group_norm(x.reshape(N, -1, C).transpose(1, 2), groups, gamma, beta, eps).transpose(1, 2).reshape(shp)
        return (F.silu(y) if silu else y).contiguous()

    def layernorm(self, x, gamma, beta, eps=1e-5, pe=None, T=0, F_=0, **kw):
        self._count("layernorm")
        F_ = kw.get("F", F_)
        C = x.shape[-1]
        y = F.layer_norm(x, (C,), gamma, beta, eps)
        if pe is not None:
            rows = y.numel() // C
            frame = (torch.arange(rows) // T) % F_
            y = (y.reshape(rows, C) + pe[frame]).reshape(x.shape)
        return y.contiguous()

    # ------------------------------------------------------------------ gemm / conv
    def row_stats(self, x, eps=1e-5):
        self._count("row_stats")
        x2 = x.reshape(-1, x.shape[-1]).float()
        mean = x2.mean(dim=1)
        var = x2.var(dim=1, unbiased=False)
        return torch.stack([mean, torch.rsqrt(var + eps)], dim=1).contiguous()

    def pad_channels(self, x, c_pad):
        out = torch.zeros(tuple(x.shape[:-1]) + (c_pad,), dtype=x.dtype)
        out[..., : x.shape[-1]] = x
        return out

    def gemm(self, A, W, bias=None, rowscale=None, rowbias=None, rows_per_group=0, residual=None, alpha=1.0, geglu_block=0,
             out=None, out_f32=False, dtype=None, exchange=None, rowbias_mod=0, rowstats=None, colsum=None, act=0):
        self._count("gemm")
        K = A.shape[-1]
        a2 = A.reshape(-1, K).float()
        y = a2 @ W.float().t()
        if rowstats is not None:
            y = rowstats[:, 1:2] * (y - rowstats[:, 0:1] * colsum[None, :])
        if bias is not None:
            y = y + bias
        if geglu_block:
            n = W.shape[0]
            y = y.view(-1, n // (2 * geglu_block), 2, geglu_block)
            y = (y[:, :, 0] * F.gelu(y[:, :, 1])).reshape(-1, n // 2)
        if rowscale is not None:
            y = y * rowscale[:, None]
        y = y * alpha
        if rowbias is not None:
            grp = torch.arange(y.shape[0]) // rows_per_group
            if rowbias_mod:
                grp = grp % rowbias_mod
            y = y + rowbias[grp]
        if act == 1:
            y = F.silu(y)
        elif act == 2:
            y = F.relu(y)
        if residual is not None:
            y = y + residual.reshape(y.shape)
        if exchange is not None:
            exchange.deliver(y)
            return None
        y = y.reshape(tuple(A.shape[:-1]) + (y.shape[-1],)).contiguous()
        if out is not None:
            out.copy_(y)
            return out
        return y

    def conv3x3(self, x, w_krsc, bias=None, rowbias=None, frames_per_group=0, residual=None, stride=1, upsample2x=False,
                w_subpixel=None, act=0):
        self._count("conv3x3")
        xin = x.float().permute(0, 3, 1, 2)
        if upsample2x and w_subpixel is not None:
            # the four 2x2-tap parity convolutions of mmgt_conv3x3 (w_subpixel layout: include/mmgt_b200.h)
            N, Cin, H, W = xin.shape
            y = torch.zeros(N, w_subpixel.shape[0], 2 * H, 2 * W)
            xp = F.pad(xin, (1, 1, 1, 1))
            for a in (0, 1):
                for b in (0, 1):
                    for ty in (0, 1):
                        for tx in (0, 1):
                            oy, ox = ty - (0 if a else 1), tx - (0 if b else 1)
                            patch = xp[:, :, 1 + oy:1 + oy + H, 1 + ox:1 + ox + W]
                            y[:, :, a::2, b::2] += torch.einsum("nchw,oc->nohw", patch, w_subpixel[:, 2 * a + b, ty, tx].float())
            if bias is not None:
                y = y + bias[None, :, None, None]
            y = y.permute(0, 2, 3, 1)
        else:
            if upsample2x:
                xin = F.interpolate(xin, scale_factor=2.0, mode="nearest")
            y = F.conv2d(xin, w_krsc.float().permute(0, 3, 1, 2), bias, stride=stride, padding=1).permute(0, 2, 3, 1)
        if rowbias is not None:
            y = y + rowbias[torch.arange(y.shape[0]) // frames_per_group][:, None, None, :]
        if act == 1:
            y = F.silu(y)
        elif act == 2:
            y = F.relu(y)
        if residual is not None:
            y = y + residual
        return y.contiguous()

    # ------------------------------------------------------------------ attention
    @staticmethod
    def _sdpa(q, k, v, heads):
        Lq, C = q.shape
        d = C // heads
        qh, kh, vh = (t.reshape(-1, heads, d).transpose(0, 1) for t in (q, k, v))
        return F.scaled_dot_product_attention(qh[None], kh[None], vh[None])[0].transpose(0, 1).reshape(Lq, C)

    def attention(self, q, k, v, heads, k2=None, v2=None, seg2_index=None, kv_batch_stride=0):
        self._count("attention")
        out = []
        for n in range(q.shape[0]):
            kk, vv = k[n], v[n]
            if k2 is not None:
                s = 0 if seg2_index is None else int(seg2_index[n])
                if s >= 0:
                    kk, vv = torch.cat([kk, k2[s]]), torch.cat([vv, v2[s]])
            out.append(self._sdpa(q[n].float(), kk.float(), vv.float(), heads))
        return torch.stack(out).contiguous()

    def audio_attention_supported(self, M, d):
        return self.fuse_audio and M <= 32 and d % 8 == 0

    def audio_attention(self, q3, kv6, masks, scale, N, T, heads):
        self._count("audio_attention")
        rows, C3 = q3.shape
        C = C3 // 3
        M = kv6.shape[0] // N
        out = torch.zeros(rows, C3 + 8)
        for r in range(3):
            gate = masks[r] * float(scale[r])
            for n in range(N):
                q = q3[n * T:(n + 1) * T, r * C:(r + 1) * C]
                k = kv6[n * M:(n + 1) * M, 2 * r * C:(2 * r + 1) * C]
                v = kv6[n * M:(n + 1) * M, (2 * r + 1) * C:(2 * r + 2) * C]
                out[n * T:(n + 1) * T, r * C:(r + 1) * C] = self._sdpa(q, k, v, heads) * gate[n * T:(n + 1) * T, None]
            out[:, 3 * C + r] = gate
        return out

    def temporal_attention(self, qkv, B, Fr, T, heads):
        self._count("temporal_attention")
        C = qkv.shape[-1] // 3
        d = C // heads
        x = qkv.reshape(B, Fr, T, 3, heads, d).permute(3, 0, 2, 4, 1, 5).reshape(3, B * T, heads, Fr, d)   # (b d) f c per head
        o = F.scaled_dot_product_attention(x[0], x[1], x[2])                                              # (B*T, heads, Fr, d)
        return o.reshape(B, T, heads, Fr, d).permute(0, 3, 1, 2, 4).reshape(B * Fr * T, C).contiguous()

    # ------------------------------------------------------------------ small pieces
    def timestep_embedding(self, t, dim, flip, shift):
        half = dim // 2
        freq = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32) / (half - shift))
        arg = t.float()[:, None] * freq[None]
        s, c = torch.sin(arg), torch.cos(arg)
        return torch.cat([c, s], dim=1) if flip else torch.cat([s, c], dim=1)

    def silu_f32(self, x):
        return F.silu(x)

    def upsample_nearest2x(self, x):
        return x.repeat_interleave(2, dim=1).repeat_interleave(2, dim=2)

    def gather_rows(self, src, idx, out=None):
        y = src[idx.long()]
        if out is not None:
            out.copy_(y.reshape(out.shape))
            return out
        return y

    def window_accumulate(self, noise_acc, pred, frames, b0):
        Bp = pred.shape[0]
        noise_acc[b0:b0 + Bp, :, frames.long()] += pred.float()

    def cfg_ddim_step(self, latents, noise_acc, inv_count, cfg, guidance, cx, cv):
        ic = inv_count.view(1, -1, 1, 1)
        u = noise_acc[0] * ic
        v = u + guidance * (noise_acc[1] * ic - u) if cfg else u
        latents[0] = cx * latents[0] + cv * v

    def row_exchange_copy(self, src, exchange):
        exchange.deliver(src)


class FakeExchange:
    """Row exchange between emulated shards on the CPU, driven by the production row mapping (frame_shard.exchange_destination)."""

    def __init__(self, group, direction, B, F_, T, C):
        self.group, self.direction, self.B, self.F, self.T, self.C = group, direction, B, F_, T, C
        rows = B * (F_ // group.k) * T
        self.recv = torch.zeros(rows, C)

    def deliver(self, y):
        from mmgt_b200.frame_shard import exchange_destination
        g = self.group
        key = (self.direction, self.B, self.F, self.T, self.C, g.epoch_of(self))
        dest = [exchange_destination(self.direction, m, g.k, g.shard, self.B, self.F, self.T) for m in range(y.shape[0])]
        for s in range(g.k):
            src = torch.tensor([m for m, (ss, _) in enumerate(dest) if ss == s], dtype=torch.long)
            rows = torch.tensor([r for ss, r in dest if ss == s], dtype=torch.long)
            g.world[s].mailbox(key)[rows] = y[src]


class FakeShardGroup:
    """k emulated shards that run in lock-step threads; `barrier()` is a real threading.Barrier."""

    def __init__(self, k, shard, world: List["FakeShardGroup"], barrier):
        self.k, self.shard, self.world, self._barrier = k, shard, world, barrier
        self._n = 0
        self._boxes = {}
        self._epochs = {}
        import threading
        self._lock = threading.Lock()      # peers and the owner may ask for the same mailbox at the same time

    def epoch_of(self, ex):
        return self._epochs[id(ex)]

    def mailbox(self, key):
        with self._lock:
            if key not in self._boxes:
                direction, B, F_, T, C, _ = key
                self._boxes[key] = torch.zeros(B * (F_ // self.k) * T, C)
            return self._boxes[key]

    def exchange(self, direction, B, F_, T, C):
        if F_ % self.k or T % self.k:
            raise ValueError(f"frame sharding needs frames ({F_}) and tokens per frame ({T}) divisible by {self.k}")
        ex = FakeExchange(self, direction, B, F_, T, C)
        self._epochs[id(ex)] = self._n
        self._n += 1
        key = (direction, B, F_, T, C, self._epochs[id(ex)])
        ex.recv = self.mailbox(key)
        return ex

    def barrier(self):
        self._barrier.wait(timeout=120)

    @classmethod
    def make(cls, k):
        import threading
        world: List[FakeShardGroup] = []
        bar = threading.Barrier(k)
        for s in range(k):
            world.append(cls(k, s, world, bar))
        return world
