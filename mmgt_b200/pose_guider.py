"""Host mirror of src/models/pose_guider.py on the sm_100a kernels.

``PoseGuider`` keeps the reference's constructor and state-dict keys (``conv_in``, ``blocks.0..5``, ``conv_out``; all
``InflatedConv3d`` 3x3) and its ``forward(conditioning (B, 3, F, H, W)) -> (B, C_emb, F, H/8, W/8)``.  Execution: the
eight convolutions run as tensor-core implicit GEMMs (``mmgt_conv3x3``: stride 1, and stride 2 through the TMA traversal
stride) with SiLU in the epilogue (``act = 1``, pose_guider.py:47-55); the narrow channel counts of the first layers
(3, 16, 32) are zero-padded to 64 IN THE WEIGHT PACKS, so activations stay 64 channels wide without extra passes
(a zero output channel stays zero through SiLU and meets zero weights in the next layer).  float32 tier: CUDA-core kernel.
"""
from typing import Tuple

import torch
import torch.nn as nn

from .kernels import Engine, get_engine
from .packing import Pack, f32
from .resnet import InflatedConv3d


def _pad64(c: int) -> int:
    return max(64, (c + 63) // 64 * 64)


class PoseGuider(nn.Module):
    def __init__(self, conditioning_embedding_channels: int, conditioning_channels: int = 3,
                 block_out_channels: Tuple[int, ...] = (16, 32, 64, 128)):
        super().__init__()
        self.conv_in = InflatedConv3d(conditioning_channels, block_out_channels[0], kernel_size=3, padding=1)
        self.blocks = nn.ModuleList([])
        for i in range(len(block_out_channels) - 1):
            cin, cout = block_out_channels[i], block_out_channels[i + 1]
            self.blocks.append(InflatedConv3d(cin, cin, kernel_size=3, padding=1))
            self.blocks.append(InflatedConv3d(cin, cout, kernel_size=3, padding=1, stride=2))
        self.conv_out = InflatedConv3d(block_out_channels[-1], conditioning_embedding_channels, kernel_size=3, padding=1)
        for p in self.conv_out.parameters():       # zero_module (pose_guider.py:38-45)
            nn.init.zeros_(p)
        self.compute_dtype = None
        self._pack = Pack()

    @property
    def dtype(self):
        return self.conv_in.weight.dtype

    @property
    def device(self):
        return self.conv_in.weight.device

    def layers(self):
        return [self.conv_in] + list(self.blocks) + [self.conv_out]

    def _packed(self, eng: Engine):
        convs = self.layers()
        bf = eng.dtype == torch.bfloat16

        def build():
            out = []
            for k, c in enumerate(convs):
                w = c.weight.detach().to(eng.device).permute(0, 2, 3, 1).float()       # (Cout, 3, 3, Cin)
                cout, cin = w.shape[0], w.shape[-1]
                last = k == len(convs) - 1
                cin_p = _pad64(cin) if bf else (cin + 3) // 4 * 4
                cout_p = cout if (last or not bf) else _pad64(cout)
                wp = torch.zeros((cout_p, 3, 3, cin_p), device=eng.device, dtype=torch.float32)
                wp[:cout, :, :, :cin] = w
                bp = torch.zeros(cout_p, device=eng.device, dtype=torch.float32)
                bp[:cout] = f32(c.bias, eng)
                out.append((wp.to(eng.dtype).contiguous(), bp, c.stride[0], cout))
            return out
        return self._pack.get(eng, [p for c in convs for p in (c.weight, c.bias)], build)

    def _engine(self, x) -> Engine:
        dt = self.compute_dtype or self.dtype
        if dt == torch.float16:
            dt = torch.bfloat16
        return get_engine(x.device, dt)

    @torch.no_grad()
    def forward(self, conditioning):
        B, C, F, H, W = conditioning.shape
        eng = self._engine(conditioning)
        packs = self._packed(eng)
        x = eng.ncfhw_to_tokens(conditioning)                                           # (B*F, H, W, 3)
        for k, (w, b, stride, cout) in enumerate(packs):
            if x.shape[-1] < w.shape[-1]:
                x = eng.pad_channels(x, w.shape[-1])
            x = eng.conv3x3(x, w, bias=b, stride=stride, act=0 if k == len(packs) - 1 else 1)
        out_dtype = conditioning.dtype if conditioning.dtype in (torch.float32, torch.bfloat16) else torch.float32
        y = eng.tokens_to_ncfhw(x, B, F, out_dtype)
        return y if y.dtype == conditioning.dtype else y.to(conditioning.dtype)
