"""Host mirror of src/models/audio_proj.py on the sm_100a kernels.

``AudioProjModel`` keeps the reference's constructor, state-dict keys (``proj1/2/3``, ``norm``) and
``forward(audio_embeds (bz, f, window, blocks, channels)) -> (bz, f, context_tokens, output_dim)``: three Linear layers
with ReLU in the GEMM epilogue (``act = 2``, audio_proj.py:108-111), reshape to tokens, LayerNorm (:117).
"""
import torch
import torch.nn as nn

from .kernels import Engine, get_engine
from .packing import Pack, f32, run


class AudioProjModel(nn.Module):
    def __init__(self, seq_len=5, blocks=12, channels=768, intermediate_dim=512, output_dim=768, context_tokens=32):
        super().__init__()
        self.seq_len, self.blocks, self.channels = seq_len, blocks, channels
        self.input_dim = seq_len * blocks * channels
        self.intermediate_dim, self.context_tokens, self.output_dim = intermediate_dim, context_tokens, output_dim
        self.proj1 = nn.Linear(self.input_dim, intermediate_dim)
        self.proj2 = nn.Linear(intermediate_dim, intermediate_dim)
        self.proj3 = nn.Linear(intermediate_dim, context_tokens * output_dim)
        self.norm = nn.LayerNorm(output_dim)
        self.compute_dtype = None
        self._pack = Pack()

    @property
    def dtype(self):
        return self.proj1.weight.dtype

    @property
    def device(self):
        return self.proj1.weight.device

    def _engine(self, x) -> Engine:
        dt = self.compute_dtype or self.dtype
        if dt == torch.float16:
            dt = torch.bfloat16
        return get_engine(x.device, dt)

    @torch.no_grad()
    def forward(self, audio_embeds):
        bz, f = audio_embeds.shape[:2]
        eng = self._engine(audio_embeds)
        lin = (self.proj1, self.proj2, self.proj3)
        pk = self._pack.get(eng, [p for m in lin for p in (m.weight, m.bias)] + [self.norm.weight, self.norm.bias],
                            lambda: [(run(m.weight, eng), f32(m.bias, eng)) for m in lin]
                            + [(f32(self.norm.weight, eng), f32(self.norm.bias, eng))])
        x = audio_embeds.reshape(bz * f, -1).to(eng.dtype).contiguous()
        x = eng.gemm(x, pk[0][0], bias=pk[0][1], act=2)
        x = eng.gemm(x, pk[1][0], bias=pk[1][1], act=2)
        t = eng.gemm(x, pk[2][0], bias=pk[2][1]).view(bz * f * self.context_tokens, self.output_dim)
        t = eng.layernorm(t, pk[3][0], pk[3][1], self.norm.eps)
        out = t.view(bz, f, self.context_tokens, self.output_dim)
        return out if out.dtype == audio_embeds.dtype else out.to(audio_embeds.dtype)
