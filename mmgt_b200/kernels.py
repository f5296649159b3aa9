"""Tensor-level wrappers over the C ABI: torch tensors in, torch tensors out, all math in libmmgt_b200.so.

torch is used only for device memory (``torch.empty``), streams and dtype bookkeeping.
Activations are channels-last token tensors ``(N, T, C)`` / ``(N, H, W, C)``, contiguous.
"""
import ctypes as C
from typing import Optional

import torch

from . import _lib
from ._lib import AttentionParams, AudioAttentionParams, Context, Conv3x3Params, GemmParams, check

_DT = {torch.float32: _lib.F32, torch.bfloat16: _lib.BF16}


def dt_code(dtype: torch.dtype) -> int:
    try:
        return _DT[dtype]
    except KeyError:
        raise TypeError(f"mmgt_b200 kernels run in float32 or bfloat16, not {dtype} "
                        "(the reference's fp16 mode maps to bfloat16 here)") from None


def _p(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class Engine:
    """All kernels for one device + one storage dtype."""

    def __init__(self, device: torch.device, dtype: torch.dtype):
        if device.type != "cuda":
            raise RuntimeError("mmgt_b200 has no CPU path: tensors must live on a CUDA (sm_100a) device")
        self.device = device
        self.dtype = dtype
        self.dt = dt_code(dtype)
        idx = device.index if device.index is not None else torch.cuda.current_device()
        self.ctx = Context.get(idx)
        self.lib = self.ctx.lib
        self.h = self.ctx.handle
        self._stats_ws = None
        self._conv_ws = None
        self.unfused_exchange = False   # A/B switch: GEMM + mmgt_row_exchange_copy instead of the fused epilogue
        # bf16 tensor-core tier: LayerNorm applied in the consuming GEMM's epilogue (row statistics kernel + rowstats /
        # colsum).  Off by default: at the model's widths the GEMMs that consume a LayerNorm are paced by their epilogue
        # (K = 320 .. 1280), and the extra epilogue work costs what the saved pass gains (479.6 vs 481.6 ms per step,
        # profiles/r2_ab_flags.md).  Kept as a tested option (bench.py --fuse-ln 1).
        self.fuse_layernorm = False
        self.prof = None          # optional: dict key -> [events..., flops, bytes] filled by bench.py's roofline pass

    # ------------------------------------------------------------------ per-launch timing (bench.py only)
    def _t0(self):
        if self.prof is None:
            return None
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        return ev

    def _t1(self, ev0, key, flops=0.0, nbytes=0.0):
        if ev0 is None:
            return
        ev1 = torch.cuda.Event(enable_timing=True)
        ev1.record()
        rec = self.prof.setdefault(key, dict(events=[], flops=0.0, bytes=0.0, calls=0))
        rec["events"].append((ev0, ev1))
        rec["flops"] += flops
        rec["bytes"] += nbytes
        rec["calls"] += 1

    # ------------------------------------------------------------------ helpers
    @property
    def ln_fused(self) -> bool:
        """LayerNorm is folded into the GEMM that consumes it (row statistics + epilogue) on the bf16 tensor-core tier."""
        return self.fuse_layernorm and self.dtype == torch.bfloat16 and self.ctx.tensor_cores()

    @property
    def subpixel_upsample(self) -> bool:
        """Upsample3D as four 2x2-tap sub-pixel convolutions (needs the pre-summed weight pack): bf16 tensor-core tier."""
        return self.dtype == torch.bfloat16 and self.ctx.tensor_cores()

    def empty(self, *shape, dtype=None):
        return torch.empty(shape, device=self.device, dtype=dtype or self.dtype)

    def _stats(self, n_doubles: int):
        if self._stats_ws is None or self._stats_ws.numel() < n_doubles:
            self._stats_ws = torch.empty(max(n_doubles, 8192), device=self.device, dtype=torch.float64)
        return self._stats_ws

    def geglu_block(self, n_rows: int) -> int:
        """Row-interleave granularity a GEGLU weight with ``n_rows`` rows should be packed with: 16 for the
        tensor-core kernel (any N-tile width then holds matching value / gate columns), else the natural halves."""
        if self.dtype == torch.bfloat16 and n_rows % 32 == 0 and self.lib.mmgt_gemm_tc_block_n(int(n_rows)):
            return 16
        return n_rows // 2

    # ------------------------------------------------------------------ layout
    def ncfhw_to_tokens(self, x: torch.Tensor, add: Optional[torch.Tensor] = None) -> torch.Tensor:
        B, Cc, F, H, W = x.shape
        x = x.contiguous()
        if x.dtype not in _DT:
            x = x.float()
        if add is not None:
            add = add.contiguous().to(x.dtype)
        out = self.empty(B * F, H, W, Cc)
        check(self.lib.mmgt_ncfhw_to_tokens(self.h, _p(x), _p(add), _p(out), B, Cc, F, H, W, dt_code(x.dtype), self.dt,
                                             _stream()), "mmgt_ncfhw_to_tokens")
        return out

    def tokens_to_ncfhw(self, x: torch.Tensor, B: int, F: int, out_dtype: torch.dtype) -> torch.Tensor:
        """x: (N, H, W, C) tokens; the channel dim may be a slice of a wider (padded) buffer."""
        N, H, W, Cc = x.shape
        ld = x.stride(2)
        assert x.stride(3) == 1 and x.stride(1) == W * ld and x.stride(0) == H * W * ld
        out = torch.empty((B, Cc, F, H, W), device=self.device, dtype=out_dtype)
        check(self.lib.mmgt_tokens_to_ncfhw(self.h, _p(x), _p(out), B, Cc, F, H, W, ld, self.dt, dt_code(out_dtype),
                                             _stream()), "mmgt_tokens_to_ncfhw")
        return out

    # ------------------------------------------------------------------ norms
    def groupnorm(self, x1, x2, gamma, beta, groups: int, eps: float, silu: bool):
        N = x1.shape[0]
        C1 = x1.shape[-1]
        C2 = x2.shape[-1] if x2 is not None else 0
        if x2 is not None and tuple(x2.shape[:-1]) != tuple(x1.shape[:-1]):
            raise ValueError(f"groupnorm: x2 {tuple(x2.shape)} must share the leading dimensions of x1 {tuple(x1.shape)}")
        T = x1.numel() // (N * C1)
        out = self.empty(*x1.shape[:-1], C1 + C2)
        ws = self._stats(2 * N * groups + (N + 1) // 2)
        ev = self._t0()
        check(self.lib.mmgt_groupnorm(self.h, _p(x1), _p(x2), _p(out), _p(gamma), _p(beta), _p(ws), N, T, C1, C2, groups,
                                       float(eps), int(silu), self.dt, _stream()), "mmgt_groupnorm")
        self._t1(ev, ("groupnorm", C1 + C2), 0.0, 3.0 * out.numel() * out.element_size())
        return out

    def layernorm(self, x, gamma, beta, eps: float = 1e-5, pe=None, T: int = 0, F: int = 0):
        Cc = x.shape[-1]
        rows = x.numel() // Cc
        out = torch.empty_like(x)
        ev = self._t0()
        check(self.lib.mmgt_layernorm(self.h, _p(x), _p(out), _p(gamma), _p(beta), _p(pe), rows, Cc, T, F, float(eps),
                                       self.dt, _stream()), "mmgt_layernorm")
        self._t1(ev, ("layernorm", Cc), 0.0, 2.0 * out.numel() * out.element_size())
        return out

    # ------------------------------------------------------------------ gemm / conv
    def gemm(self, A, W, bias=None, rowscale=None, rowbias=None, rows_per_group: int = 0, residual=None,
             alpha: float = 1.0, geglu_block: int = 0, out=None, out_f32: bool = False, dtype=None, exchange=None,
             rowbias_mod: int = 0, rowstats=None, colsum=None, act: int = 0):
        """D = act(alpha * rowscale * (A @ W^T + bias) + rowbias[row // rows_per_group]) + residual (then GEGLU).
        A: (..., K) rows with an arbitrary leading stride on the last-but-one dim; W: (N, K).
        ``rowbias`` may be a column slice of a wider float32 matrix (its row stride is passed on); ``rowbias_mod`` wraps
        the group index.  ``rowstats`` (rows, 2) + ``colsum`` (N): the LayerNorm feeding this GEMM is applied in the
        epilogue (W must then be W diag(gamma), bias = bias + W beta).  ``act``: 0 none, 1 SiLU, 2 ReLU.
        ``exchange``: a frame_shard.Exchange -- the result rows go to the shards that own them (peer stores from
        the GEMM epilogue on the tensor-core path, GEMM + row-exchange copy otherwise); returns None."""
        K = A.shape[-1]
        M = A.numel() // K if A.is_contiguous() else A.shape[0]
        lda = K if A.is_contiguous() else A.stride(-2)
        N = W.shape[0]
        ldw = W.stride(0)
        n_out = N // 2 if geglu_block else N
        dt = self.dt if dtype is None else dt_code(dtype)
        fused_exchange = exchange is not None and self.dtype == torch.bfloat16 and self.ctx.tensor_cores() \
            and not self.unfused_exchange and self.lib.mmgt_gemm_tc_block_n(int(N)) > 0 and K % 8 == 0 and n_out % 16 == 0
        if exchange is not None and not fused_exchange:
            tmp = self.gemm(A, W, bias=bias, rowscale=rowscale, rowbias=rowbias, rows_per_group=rows_per_group,
                            residual=residual, alpha=alpha, geglu_block=geglu_block, rowbias_mod=rowbias_mod,
                            rowstats=rowstats, colsum=colsum, act=act)
            self.row_exchange_copy(tmp.view(-1, n_out), exchange)
            return None
        if fused_exchange:
            out = exchange.recv_dummy
        if out is None:
            odt = torch.float32 if out_f32 else (self.dtype if dtype is None else dtype)
            out = torch.empty(tuple(A.shape[:-1]) + (n_out,), device=self.device, dtype=odt)
        p = GemmParams()
        p.A, p.W, p.D = A.data_ptr(), W.data_ptr(), out.data_ptr()
        p.exchange = C.addressof(exchange.desc) if fused_exchange else None
        p.bias = bias.data_ptr() if bias is not None else None
        p.rowscale = rowscale.data_ptr() if rowscale is not None else None
        p.rowbias = rowbias.data_ptr() if rowbias is not None else None
        p.residual = residual.data_ptr() if residual is not None else None
        p.lda, p.ldw, p.ldd = lda, ldw, out.stride(-2) if out.dim() > 1 else n_out
        p.ldr = residual.stride(-2) if residual is not None else 0
        p.M, p.N, p.K = M, N, K
        p.rows_per_group = rows_per_group
        p.alpha = alpha
        p.geglu_block = geglu_block
        p.dtype = dt
        p.out_f32 = int(out_f32)
        if rowbias is not None:
            assert rowbias.dtype == torch.float32 and rowbias.stride(-1) == 1
            p.ld_rowbias = rowbias.stride(0) if rowbias.dim() > 1 else 0
        p.rowbias_mod = rowbias_mod
        p.rowstats = rowstats.data_ptr() if rowstats is not None else None
        p.colsum = colsum.data_ptr() if colsum is not None else None
        p.act = act
        ev = self._t0()
        check(self.lib.mmgt_gemm(self.h, C.byref(p), _stream()), "mmgt_gemm")
        self._t1(ev, ("gemm_exchange" if fused_exchange else "gemm", N, K, bool(geglu_block)), 2.0 * M * N * K,
                 (M * K + N * K + M * n_out * (2 if residual is not None else 1)) * A.element_size())
        return None if fused_exchange else out

    def row_exchange_copy(self, src, exchange):
        """Stand-alone form of the row exchange: src (rows, C) -> the owning shards' receive buffers."""
        rows, Cc = src.shape
        ev = self._t0()
        check(self.lib.mmgt_row_exchange_copy(self.h, _p(src), src.stride(0), Cc, dt_code(src.dtype), C.byref(exchange.desc),
                                               _stream()), "mmgt_row_exchange_copy")
        self._t1(ev, ("row_exchange_copy", Cc), 0.0, 2.0 * src.numel() * src.element_size())

    def row_stats(self, x, eps: float = 1e-5):
        """(mean, rstd) per row of x (rows, C) -> (rows, 2) float32: the LayerNorm statistics a GEMM applies in its
        epilogue (``gemm(..., rowstats=, colsum=)``)."""
        Cc = x.shape[-1]
        rows = x.numel() // Cc if x.is_contiguous() else x.shape[0]
        ld = Cc if x.is_contiguous() else x.stride(-2)
        out = torch.empty((rows, 2), device=self.device, dtype=torch.float32)
        ev = self._t0()
        check(self.lib.mmgt_row_stats(self.h, _p(x), _p(out), rows, Cc, ld, float(eps), dt_code(x.dtype), _stream()),
              "mmgt_row_stats")
        self._t1(ev, ("row_stats", Cc), 0.0, float(rows) * Cc * x.element_size())
        return out

    def conv3x3(self, x, w_krsc, bias=None, rowbias=None, frames_per_group: int = 0, residual=None, stride: int = 1,
                upsample2x: bool = False, w_subpixel=None, act: int = 0):
        N, H, W, Cin = x.shape
        Cout = w_krsc.shape[0]
        Hi, Wi = (2 * H, 2 * W) if upsample2x else (H, W)
        Ho, Wo = (Hi - 1) // stride + 1, (Wi - 1) // stride + 1
        out = self.empty(N, Ho, Wo, Cout)
        p = Conv3x3Params()
        p.x, p.w, p.y = x.data_ptr(), w_krsc.data_ptr(), out.data_ptr()
        p.bias = bias.data_ptr() if bias is not None else None
        p.rowbias = rowbias.data_ptr() if rowbias is not None else None
        p.residual = residual.data_ptr() if residual is not None else None
        p.N, p.H, p.W, p.Cin, p.Cout = N, H, W, Cin, Cout
        p.stride, p.upsample2x, p.frames_per_group, p.dtype = stride, int(upsample2x), frames_per_group, self.dt
        if rowbias is not None:
            assert rowbias.dtype == torch.float32 and rowbias.stride(-1) == 1
            p.ld_rowbias = rowbias.stride(0) if rowbias.dim() > 1 else 0
        p.act = act
        p.w_subpixel = w_subpixel.data_ptr() if (w_subpixel is not None and upsample2x) else None
        need = self.lib.mmgt_conv3x3_workspace_bytes(self.h, C.byref(p))
        ws = None
        if need > 0:
            if self._conv_ws is None or self._conv_ws.numel() < need:
                self._conv_ws = torch.empty(need, device=self.device, dtype=torch.uint8)
            ws = self._conv_ws
        ev = self._t0()
        check(self.lib.mmgt_conv3x3(self.h, C.byref(p), _p(ws), need if need > 0 else 0, _stream()), "mmgt_conv3x3")
        self._t1(ev, ("conv3x3", Cin, Cout, stride, int(upsample2x)), 2.0 * N * Ho * Wo * 9 * Cin * Cout,
                 (x.numel() + w_krsc.numel() + out.numel() * (2 if residual is not None else 1)) * x.element_size())
        return out

    # ------------------------------------------------------------------ attention
    def attention(self, q, k, v, heads: int, k2=None, v2=None, seg2_index=None, kv_batch_stride: int = 0):
        """q: (N, Lq, C) view (last dim contiguous); k, v: (N, Lk, C) views; k2, v2: (B2, Lk2, C) views."""
        N, Lq, Cc = q.shape
        d = Cc // heads
        out = self.empty(N, Lq, Cc)
        p = AttentionParams()
        p.q, p.k, p.v, p.out = q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr()
        p.ldq, p.ldk, p.ldv, p.ldo = q.stride(1), k.stride(1), v.stride(1), Cc
        p.kv_batch_stride = kv_batch_stride if kv_batch_stride else k.stride(0)
        p.N, p.Lq, p.Lk, p.heads, p.d = N, Lq, k.shape[1], heads, d
        if k2 is not None:
            p.k2, p.v2 = k2.data_ptr(), v2.data_ptr()
            p.ldk2, p.ldv2, p.Lk2, p.B2 = k2.stride(1), v2.stride(1), k2.shape[1], k2.shape[0]
            p.seg2_index = seg2_index.data_ptr() if seg2_index is not None else None
        p.scale = float(d) ** -0.5
        p.dtype = self.dt
        ev = self._t0()
        check(self.lib.mmgt_attention(self.h, C.byref(p), _stream()), "mmgt_attention")
        # algorithmic work: frames whose seg2 index is -1 (the CFG uncond half) attend to the first segment only
        n2 = 0 if k2 is None else (N if seg2_index is None else getattr(seg2_index, "_n_seg2", N))
        keys = N * k.shape[1] + (n2 * k2.shape[1] if k2 is not None else 0)        # sum over frames of their key count
        self._t1(ev, ("attention", d, k.shape[1], k2 is not None), 4.0 * heads * Lq * keys * d,
                 (2 * q.numel() + 2 * keys * Cc) * q.element_size())
        return out

    def audio_attention_supported(self, M: int, d: int) -> bool:
        """The fused three-region MM-HAA kernel: bf16 tensor-core tier, <= 32 audio tokens per frame."""
        return self.dtype == torch.bfloat16 and self.ctx.tensor_cores() and M <= 32 and d % 8 == 0 and d <= 160

    def audio_attention(self, q3, kv6, masks, scale, N: int, T: int, heads: int):
        """q3 (N*T, 3C), kv6 (N*M, 6C), masks 3 x (N*T,) float32, scale 3 floats -> (N*T, 3C + 8) gated outputs of the
        three regions + the gate columns (layout: mmgt_audio_attention in include/mmgt_b200.h)."""
        rows, C3 = q3.shape
        Cc = C3 // 3
        d = Cc // heads
        M = kv6.shape[0] // N
        out = self.empty(rows, C3 + 8)
        p = AudioAttentionParams()
        p.q3, p.kv6, p.out = q3.data_ptr(), kv6.data_ptr(), out.data_ptr()
        for r in range(3):
            p.mask[r] = masks[r].data_ptr()
            p.scale[r] = float(scale[r])
        p.ldq, p.ldkv, p.ldo = q3.stride(0), kv6.stride(0), out.stride(0)
        p.N, p.T, p.M, p.heads, p.d = N, T, M, heads, d
        p.softmax_scale = float(d) ** -0.5
        p.dtype = self.dt
        ev = self._t0()
        check(self.lib.mmgt_audio_attention(self.h, C.byref(p), _stream()), "mmgt_audio_attention")
        self._t1(ev, ("audio_attention", d, M), 4.0 * 3 * rows * heads * M * d, (q3.numel() + out.numel()) * q3.element_size())
        return out

    def temporal_attention(self, qkv, B: int, F: int, T: int, heads: int):
        Cc = qkv.shape[-1] // 3
        d = Cc // heads
        out = self.empty(B * F * T, Cc)
        ev = self._t0()
        check(self.lib.mmgt_temporal_attention(self.h, _p(qkv), _p(out), B, F, T, heads, d, float(d) ** -0.5, self.dt,
                                                _stream()), "mmgt_temporal_attention")
        self._t1(ev, ("temporal_attention", d), 4.0 * B * T * heads * F * F * d, (qkv.numel() + out.numel()) * out.element_size())
        return out

    # ------------------------------------------------------------------ small pieces
    def timestep_embedding(self, t_f32, dim: int, flip: bool, shift: float):
        B = t_f32.numel()
        out = torch.empty((B, dim), device=self.device, dtype=torch.float32)
        check(self.lib.mmgt_timestep_embedding(self.h, _p(t_f32), _p(out), B, dim, int(flip), float(shift), _stream()),
              "mmgt_timestep_embedding")
        return out

    def silu_f32(self, x):
        out = torch.empty_like(x)
        check(self.lib.mmgt_silu_f32(self.h, _p(x), _p(out), x.numel(), _stream()), "mmgt_silu_f32")
        return out

    def upsample_nearest2x(self, x):
        N, H, W, Cc = x.shape
        out = self.empty(N, 2 * H, 2 * W, Cc)
        check(self.lib.mmgt_upsample_nearest2x(self.h, _p(x), _p(out), N, H, W, Cc, self.dt, _stream()),
              "mmgt_upsample_nearest2x")
        return out

    def pad_channels(self, x, c_pad: int):
        """(..., C) -> (..., c_pad), zero-filled."""
        Cc = x.shape[-1]
        x = x.contiguous()
        out = torch.empty(tuple(x.shape[:-1]) + (c_pad,), device=self.device, dtype=x.dtype)
        check(self.lib.mmgt_pad_channels(self.h, _p(x), _p(out), x.numel() // Cc, Cc, c_pad, dt_code(x.dtype), _stream()),
              "mmgt_pad_channels")
        return out

    def gather_rows(self, src, idx_i32, out=None):
        """out[i] = src[idx[i]] along dim 0 (rows must be multiples of 16 bytes)."""
        n_out = idx_i32.numel()
        row_bytes = src[0].numel() * src.element_size()
        if out is None:
            out = torch.empty((n_out,) + tuple(src.shape[1:]), device=self.device, dtype=src.dtype)
        check(self.lib.mmgt_gather_rows(self.h, _p(src), _p(idx_i32), _p(out), n_out, row_bytes, _stream()), "mmgt_gather_rows")
        return out

    def window_accumulate(self, noise_acc, pred, frames_i32, b0: int):
        Bp, Cc, Fw, H, W = pred.shape
        L = noise_acc.shape[2]
        check(self.lib.mmgt_window_accumulate(self.h, _p(noise_acc), _p(pred), _p(frames_i32), Bp, b0, Cc, L, Fw, H * W,
                                               dt_code(pred.dtype), _stream()), "mmgt_window_accumulate")

    def cfg_ddim_step(self, latents, noise_acc, inv_count, cfg: bool, guidance: float, cx: float, cv: float):
        _, Cc, L, H, W = latents.shape
        check(self.lib.mmgt_cfg_ddim_step(self.h, _p(latents), _p(noise_acc), _p(inv_count), Cc, L, H * W, int(cfg),
                                           float(guidance), float(cx), float(cv), _stream()), "mmgt_cfg_ddim_step")

    def mask_resize(self, src_u8, S: int, offset: float = 0.0, want_u8: bool = False):
        L, Hs, Ws = src_u8.shape
        tmp = torch.empty((L, Hs, S), device=self.device, dtype=torch.uint8)
        out_f = torch.empty((L, S * S), device=self.device, dtype=torch.float32)
        out_u = torch.empty((L, S, S), device=self.device, dtype=torch.uint8) if want_u8 else None
        check(self.lib.mmgt_mask_resize(self.h, _p(src_u8), _p(tmp), _p(out_u), _p(out_f), L, Hs, Ws, S, float(offset),
                                         _stream()), "mmgt_mask_resize")
        return (out_f, out_u) if want_u8 else out_f


_engines = {}


def get_engine(device: torch.device, dtype: torch.dtype) -> Engine:
    key = (device.type, device.index if device.index is not None else torch.cuda.current_device(), dtype)
    if key not in _engines:
        _engines[key] = Engine(torch.device(key[0], key[1]), dtype)
    return _engines[key]
