"""Host mirror of src/models/transformer_3d.py: GroupNorm(1e-6) -> 1x1 proj_in -> block -> 1x1 proj_out -> +x."""
import torch
import torch.nn as nn

from .attention import AudioTemporalBasicTransformerBlock, TemporalBasicTransformerBlock
from .kernels import Engine
from .packing import Pack, conv1x1, f32
from .resnet import GroupNorm2d


class Transformer3DModel(nn.Module):
    def __init__(self, num_attention_heads=16, attention_head_dim=88, in_channels=None, num_layers=1, dropout=0.0,
                 norm_num_groups=32, cross_attention_dim=None, attention_bias=False, activation_fn="geglu",
                 num_embeds_ada_norm=None, use_linear_projection=False, only_cross_attention=False,
                 upcast_attention=False, unet_use_cross_frame_attention=None, unet_use_temporal_attention=None,
                 name=None, use_audio_module=False, depth=0, unet_block_name=None, stack_enable_blocks_name=None,
                 stack_enable_blocks_depth=None):
        super().__init__()
        if use_linear_projection:
            raise NotImplementedError("use_linear_projection=True (SD-1.5 uses 1x1 convs, transformer_3d.py:72-74)")
        if num_layers != 1:
            raise NotImplementedError("num_layers != 1")
        self.num_attention_heads, self.attention_head_dim = num_attention_heads, attention_head_dim
        inner = num_attention_heads * attention_head_dim
        self.in_channels, self.inner_dim, self.use_audio_module, self.name = in_channels, inner, use_audio_module, name
        self.norm = GroupNorm2d(num_groups=norm_num_groups, num_channels=in_channels, eps=1e-6, affine=True)
        self.proj_in = nn.Conv2d(in_channels, inner, kernel_size=1, stride=1, padding=0)
        common = dict(dropout=dropout, cross_attention_dim=cross_attention_dim, activation_fn=activation_fn,
                      num_embeds_ada_norm=num_embeds_ada_norm, attention_bias=attention_bias,
                      only_cross_attention=only_cross_attention, upcast_attention=upcast_attention,
                      unet_use_cross_frame_attention=unet_use_cross_frame_attention,
                      unet_use_temporal_attention=unet_use_temporal_attention)
        if use_audio_module:
            blk = AudioTemporalBasicTransformerBlock(inner, num_attention_heads, attention_head_dim, depth=depth,
                                                     unet_block_name=unet_block_name,
                                                     stack_enable_blocks_name=stack_enable_blocks_name,
                                                     stack_enable_blocks_depth=stack_enable_blocks_depth, **common)
        else:
            blk = TemporalBasicTransformerBlock(inner, num_attention_heads, attention_head_dim,
                                                name=f"{name}_0_TransformerBlock" if name else None, **common)
        self.transformer_blocks = nn.ModuleList([blk])
        self.proj_out = nn.Conv2d(inner, in_channels, kernel_size=1, stride=1, padding=0)
        self.gradient_checkpointing = False
        self._pack = Pack()

    def _packed(self, eng: Engine):
        return self._pack.get(eng, [self.proj_in.weight, self.proj_in.bias, self.proj_out.weight, self.proj_out.bias],
                              lambda: (conv1x1(self.proj_in.weight, eng), f32(self.proj_in.bias, eng),
                                       conv1x1(self.proj_out.weight, eng), f32(self.proj_out.bias, eng)))

    def run(self, eng: Engine, x, frames: int, **block_kwargs):
        """x: (N, H, W, C) -> same (transformer_3d.py:139-268)."""
        N, H, W, C = x.shape
        T = H * W
        wi, bi, wo, bo = self._packed(eng)
        h = self.norm.run(eng, x, None, silu=False)
        tok = eng.gemm(h.view(N * T, C), wi, bias=bi).view(N, T, self.inner_dim)
        tok = self.transformer_blocks[0].run(eng, tok, frames=frames, **block_kwargs) \
            if not self.use_audio_module else self.transformer_blocks[0].run(eng, tok, **block_kwargs)
        out = eng.gemm(tok.view(N * T, self.inner_dim), wo, bias=bo, residual=x.view(N * T, C))
        return out.view(N, H, W, C)
