"""A small AutoencoderKL-shaped decoder for tests (TEST INFRASTRUCTURE): post_quant_conv, conv_in, a mid block, up blocks of
resnets + nearest-x2 upsampling convs, GroupNorm + SiLU + conv_out, with the ``.decode(z).sample`` / ``.encode(x).latent_dist.mean``
/ ``.config.block_out_channels`` / ``.dtype`` / ``.device`` surface Pose2VideoPipeline touches (pipeline_pose2vid_long.py:70,
112-125, 424-430).  Random-init, reduced width: the real SD VAE weights are not available offline; what the tests need is a
fixed non-linear, 8x upsampling decoder to measure how latent differences show up in DECODED frames."""
import torch
import torch.nn as nn
import torch.nn.functional as F


class _Out:
    def __init__(self, **kw):
        self.__dict__.update(kw)


class _Res(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.n1, self.c1 = nn.GroupNorm(8, cin), nn.Conv2d(cin, cout, 3, padding=1)
        self.n2, self.c2 = nn.GroupNorm(8, cout), nn.Conv2d(cout, cout, 3, padding=1)
        self.skip = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x):
        h = self.c1(F.silu(self.n1(x)))
        h = self.c2(F.silu(self.n2(h)))
        return h + (x if self.skip is None else self.skip(x))


class VaeStub(nn.Module):
    def __init__(self, widths=(64, 64, 32, 16), seed=0):
        super().__init__()
        torch.manual_seed(seed)
        self.config = _Out(block_out_channels=[128, 256, 512, 512])           # -> vae_scale_factor 8
        self.post_quant_conv = nn.Conv2d(4, 4, 1)
        self.conv_in = nn.Conv2d(4, widths[0], 3, padding=1)
        self.mid = _Res(widths[0], widths[0])
        ups, c = [], widths[0]
        for i, w in enumerate(widths):
            ups.append(_Res(c, w))
            c = w
            if i < 3:
                ups.append(nn.Conv2d(c, c, 3, padding=1))                      # after a nearest x2
        self.ups = nn.ModuleList(ups)
        self.norm_out, self.conv_out = nn.GroupNorm(8, c), nn.Conv2d(c, 3, 3, padding=1)
        self.enc = nn.Conv2d(3, 4, 8, stride=8)
        self.decode_calls = []

    @property
    def dtype(self):
        return self.conv_in.weight.dtype

    @property
    def device(self):
        return self.conv_in.weight.device

    def encode(self, x):
        return _Out(latent_dist=_Out(mean=self.enc(x)))

    def decode(self, z):
        self.decode_calls.append(z.shape[0])
        h = self.mid(self.conv_in(self.post_quant_conv(z)))
        for m in self.ups:
            if isinstance(m, _Res):
                h = m(h)
            else:
                h = m(F.interpolate(h, scale_factor=2.0, mode="nearest"))
        return _Out(sample=torch.tanh(self.conv_out(F.silu(self.norm_out(h)))))
