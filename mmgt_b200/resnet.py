"""Host mirror of src/models/resnet.py: same class names / parameters / state-dict keys, execution on the
sm_100a kernels over channels-last frames ``(N = B*F, H, W, C)``.

``run(...)`` methods are the fast path used by ``UNet3DConditionModel``; the nn.Module ``forward`` of the
layer classes keeps the reference's 5-D ``(B, C, F, H, W)`` calling convention for drop-in use.
"""
from typing import Optional

import torch
import torch.nn as nn

from .kernels import Engine, get_engine
from .packing import Pack, conv1x1, conv_krsc, f32, subpixel_pack


def _engine_for(module: nn.Module, x: torch.Tensor) -> Engine:
    dt = getattr(module, "compute_dtype", None) or x.dtype
    if dt == torch.float16:
        dt = torch.bfloat16
    return get_engine(x.device, dt)


class InflatedConv3d(nn.Conv2d):
    """Per-frame Conv2d (resnet.py:9-17).  3x3 (stride 1/2, pad 1) and 1x1 kernels are supported."""

    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self._pack = Pack()
        self._sp_pack = Pack()

    def packed(self, eng: Engine):
        def build():
            b = f32(self.bias, eng) if self.bias is not None else None
            if self.kernel_size == (3, 3):
                w = conv_krsc(self.weight, eng)
                cout = w.shape[0]
                if eng.dtype == torch.bfloat16 and w.shape[-1] < 64:
                    # conv_in (4 -> 320): zero-pad the input channels to 64 so the tensor-core implicit GEMM takes it
                    # (run() pads the activation to match)
                    wp = torch.zeros(tuple(w.shape[:-1]) + (64,), device=w.device, dtype=w.dtype)
                    wp[..., : w.shape[-1]] = w
                    w = wp
                if eng.dtype == torch.bfloat16 and cout < 32 and w.shape[-1] % 64 == 0:
                    # conv_out (320 -> 4): pad the output channels to 32 so the tensor-core implicit GEMM takes it
                    wp = torch.zeros((32,) + tuple(w.shape[1:]), device=w.device, dtype=w.dtype)
                    wp[:cout] = w
                    w = wp
                    if b is not None:
                        bp = torch.zeros(32, device=b.device, dtype=b.dtype)
                        bp[:cout] = b
                        b = bp
            else:
                w = conv1x1(self.weight, eng)
            return w, b
        return self._pack.get(eng, [self.weight] + ([self.bias] if self.bias is not None else []), build)

    def subpixel(self, eng: Engine):
        """Pre-summed 2x2 taps of the four output parities for the x2-upsampling form (Upsample3D), bf16 tier only."""
        if not eng.subpixel_upsample:
            return None
        return self._sp_pack.get(eng, [self.weight], lambda: subpixel_pack(self.weight, eng))

    def run(self, eng: Engine, x, rowbias=None, frames_per_group=0, residual=None, upsample2x=False, act=0):
        w, b = self.packed(eng)
        if self.kernel_size == (3, 3):
            if self.padding != (1, 1):
                raise NotImplementedError("InflatedConv3d: only padding=1 is implemented for 3x3 kernels")
            if x.shape[-1] < w.shape[-1]:
                x = eng.pad_channels(x, w.shape[-1])
            y = eng.conv3x3(x, w, bias=b, rowbias=rowbias, frames_per_group=frames_per_group, residual=residual,
                            stride=self.stride[0], upsample2x=upsample2x, act=act,
                            w_subpixel=self.subpixel(eng) if upsample2x else None)
            # padded output channels: return the real ones as a strided view (consumers take a channel stride)
            return y if y.shape[-1] == self.out_channels else y[..., : self.out_channels]
        if self.kernel_size != (1, 1) or self.stride != (1, 1):
            raise NotImplementedError("InflatedConv3d: only 3x3 and 1x1 kernels are implemented")
        N, H, W, C = x.shape
        out = eng.gemm(x.view(N * H * W, C), w, bias=b, residual=None if residual is None else residual.view(N * H * W, -1),
                       act=act)
        return out.view(N, H, W, -1)

    def forward(self, x):
        B, C, F, H, W = x.shape
        eng = _engine_for(self, x)
        y = self.run(eng, eng.ncfhw_to_tokens(x))
        return eng.tokens_to_ncfhw(y, B, F, x.dtype)


class InflatedGroupNorm(nn.GroupNorm):
    """Per-frame GroupNorm (resnet.py:20-28)."""

    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self._pack = Pack()

    def packed(self, eng: Engine):
        return self._pack.get(eng, [self.weight, self.bias], lambda: (f32(self.weight, eng), f32(self.bias, eng)))

    def run(self, eng: Engine, x1, x2=None, silu=False):
        g, b = self.packed(eng)
        return eng.groupnorm(x1, x2, g, b, self.num_groups, self.eps, silu)

    def forward(self, x):
        B, C, F, H, W = x.shape
        eng = _engine_for(self, x)
        return eng.tokens_to_ncfhw(self.run(eng, eng.ncfhw_to_tokens(x)), B, F, x.dtype)


class GroupNorm2d(InflatedGroupNorm):
    """nn.GroupNorm used inside Transformer3DModel / the motion module (already per frame there)."""


class Upsample3D(nn.Module):
    """Nearest x2 + 3x3 conv (resnet.py:31-90); the upsample is folded into the conv's gather."""

    def __init__(self, channels, use_conv=False, use_conv_transpose=False, out_channels=None, name="conv"):
        super().__init__()
        if use_conv_transpose or not use_conv:
            raise NotImplementedError
        self.channels = channels
        self.out_channels = out_channels or channels
        self.conv = InflatedConv3d(self.channels, self.out_channels, 3, padding=1)

    def run(self, eng: Engine, x):
        assert x.shape[-1] == self.channels
        return self.conv.run(eng, x, upsample2x=True)


class Downsample3D(nn.Module):
    """Stride-2 3x3 conv (resnet.py:93-120)."""

    def __init__(self, channels, use_conv=False, out_channels=None, padding=1, name="conv"):
        super().__init__()
        if not use_conv or padding != 1:
            raise NotImplementedError
        self.channels = channels
        self.out_channels = out_channels or channels
        self.conv = InflatedConv3d(self.channels, self.out_channels, 3, stride=2, padding=padding)

    def run(self, eng: Engine, x):
        assert x.shape[-1] == self.channels
        return self.conv.run(eng, x)


class ResnetBlock3D(nn.Module):
    """GN-SiLU-conv-(+temb)-GN-SiLU-conv-(+shortcut) (resnet.py:123-247)."""

    def __init__(self, *, in_channels, out_channels=None, conv_shortcut=False, dropout=0.0, temb_channels=512, groups=32,
                 groups_out=None, pre_norm=True, eps=1e-6, non_linearity="swish", time_embedding_norm="default",
                 output_scale_factor=1.0, use_in_shortcut=None, use_inflated_groupnorm=None):
        super().__init__()
        if time_embedding_norm != "default" or non_linearity not in ("swish", "silu"):
            raise NotImplementedError("ResnetBlock3D: only default time embedding + SiLU are on the hot path")
        if output_scale_factor != 1.0:
            raise NotImplementedError("output_scale_factor != 1")
        out_channels = in_channels if out_channels is None else out_channels
        self.in_channels, self.out_channels = in_channels, out_channels
        groups_out = groups if groups_out is None else groups_out
        self.norm1 = InflatedGroupNorm(num_groups=groups, num_channels=in_channels, eps=eps, affine=True)
        self.conv1 = InflatedConv3d(in_channels, out_channels, kernel_size=3, stride=1, padding=1)
        self.time_emb_proj = nn.Linear(temb_channels, out_channels) if temb_channels is not None else None
        self.norm2 = InflatedGroupNorm(num_groups=groups_out, num_channels=out_channels, eps=eps, affine=True)
        self.conv2 = InflatedConv3d(out_channels, out_channels, kernel_size=3, stride=1, padding=1)
        use_in_shortcut = (in_channels != out_channels) if use_in_shortcut is None else use_in_shortcut
        self.conv_shortcut = InflatedConv3d(in_channels, out_channels, kernel_size=1, stride=1, padding=0) \
            if use_in_shortcut else None
        self._tpack = Pack()

    def run(self, eng: Engine, x1, x2: Optional[torch.Tensor], temb, frames: int):
        """x1 (+ optional skip x2, the virtual channel concat of unet_3d_blocks.py:894) -> (N,H,W,Cout).
        temb: TimeProjections (unet_3d.py) -- this block's time_emb_proj(silu(emb)) is a column slice of ONE projection
        computed per DDIM step for all 22 resnets -- or a (B, temb_channels) float32 tensor silu(emb)."""
        N, H, W, C1 = x1.shape
        rows = N * H * W
        if x2 is not None and tuple(x2.shape[:-1]) != (N, H, W):
            raise ValueError(f"skip connection {tuple(x2.shape)} does not match the hidden state {tuple(x1.shape)} "
                             "(latent sizes must be divisible by 2**num_upsamplers)")
        if torch.is_tensor(temb):
            wt, bt = self._tpack.get(eng, [self.time_emb_proj.weight, self.time_emb_proj.bias],
                                     lambda: (f32(self.time_emb_proj.weight, eng), f32(self.time_emb_proj.bias, eng)))
            tproj = eng.gemm(temb, wt, bias=bt, dtype=torch.float32)          # (B, Cout) fp32
        else:
            tproj = temb.of(self)                                             # (B, Cout) fp32 view, row stride = total width
        h = self.norm1.run(eng, x1, x2, silu=True)
        h = self.conv1.run(eng, h, rowbias=tproj, frames_per_group=frames)
        h = self.norm2.run(eng, h, None, silu=True)
        if self.conv_shortcut is not None:
            ws, bs = self.conv_shortcut.packed(eng)
            res = eng.gemm(x1.view(rows, C1), ws[:, :C1], bias=bs)
            if x2 is not None:
                res = eng.gemm(x2.view(rows, -1), ws[:, C1:], residual=res, out=res)
            res = res.view(N, H, W, -1)
        else:
            assert x2 is None
            res = x1
        return self.conv2.run(eng, h, residual=res)
