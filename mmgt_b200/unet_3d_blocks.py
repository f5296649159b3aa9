"""Host mirror of src/models/unet_3d_blocks.py: block sequencing, skip connections, down / up sampling.

Blocks exchange channels-last frame tensors (N, H, W, C).  Skip connections are NOT concatenated: the
(hidden, skip) pair goes to ResnetBlock3D as a virtual concat (GroupNorm reads both, the 1x1 shortcut
runs as two accumulating GEMMs), which removes the torch.cat copies of unet_3d_blocks.py:894,1057.
"""
from dataclasses import dataclass
from typing import List, Optional

import torch
import torch.nn as nn

from .kernels import Engine
from .motion_module import get_motion_module
from .resnet import Downsample3D, ResnetBlock3D, Upsample3D
from .transformer_3d import Transformer3DModel


@dataclass
class StepInputs:
    """Per-forward conditioning shared by every block (built once in UNet3DConditionModel.forward)."""
    frames: int
    temb_silu: torch.Tensor                 # (B, 1280) float32 = silu(time embedding)
    clip: torch.Tensor                      # (B, L_clip, 768)
    seg2_index: Optional[torch.Tensor]      # (N,) int32: reference-bank row per frame, -1 = self-attention only
    audio_rows: Optional[torch.Tensor]      # (N*M, 768) run dtype
    masks: Optional[list]                   # [level] -> (full, face, lip) each (N*T_level,) float32
    scale: tuple = (1.0, 1.0, 1.0)          # motion_scale as seen by MM-HAA on this call
    shard: Optional[object] = None          # frame_shard.FrameShardGroup: ``frames`` are then this rank's F/k frames


def _resnet(in_c, out_c, temb_c, eps, groups):
    return ResnetBlock3D(in_channels=in_c, out_channels=out_c, temb_channels=temb_c, eps=eps, groups=groups,
                         use_inflated_groupnorm=True)


def _spatial(heads, channels, cross_dim, groups, name=None):
    return Transformer3DModel(heads, channels // heads, in_channels=channels, num_layers=1, cross_attention_dim=cross_dim,
                              norm_num_groups=groups, unet_use_cross_frame_attention=False,
                              unet_use_temporal_attention=False, name=name)


class CrossAttnDownBlock3D(nn.Module):
    def __init__(self, in_channels, out_channels, temb_channels, num_layers=1, resnet_eps=1e-6, resnet_groups=32,
                 attn_num_head_channels=1, cross_attention_dim=1280, audio_attention_dim=1024, add_downsample=True,
                 use_motion_module=None, use_audio_module=None, depth=0, stack_enable_blocks_name=None,
                 stack_enable_blocks_depth=None, motion_module_type=None, motion_module_kwargs=None, name=None):
        super().__init__()
        self.has_cross_attention = True
        self.depth = depth
        resnets, attentions, audio_modules, motion_modules = [], [], [], []
        for i in range(num_layers):
            in_c = in_channels if i == 0 else out_channels
            resnets.append(_resnet(in_c, out_channels, temb_channels, resnet_eps, resnet_groups))
            attentions.append(_spatial(attn_num_head_channels, out_channels, cross_attention_dim, resnet_groups,
                                       name=f"{name}_{i}_TransformerModel" if name else None))
            # NB: head_dim from the layer's *input* width (unet_3d_blocks.py:467-471) => inner_dim quirk
            audio_modules.append(Transformer3DModel(
                attn_num_head_channels, in_c // attn_num_head_channels, in_channels=out_channels, num_layers=1,
                cross_attention_dim=audio_attention_dim, norm_num_groups=resnet_groups, use_audio_module=True,
                depth=depth, unet_block_name="down", stack_enable_blocks_name=stack_enable_blocks_name,
                stack_enable_blocks_depth=stack_enable_blocks_depth, unet_use_cross_frame_attention=False,
                unet_use_temporal_attention=False) if use_audio_module else None)
            motion_modules.append(get_motion_module(out_channels, motion_module_type, motion_module_kwargs)
                                  if use_motion_module else None)
        self.attentions = nn.ModuleList(attentions)
        self.resnets = nn.ModuleList(resnets)
        self.audio_modules = nn.ModuleList(audio_modules)
        self.motion_modules = nn.ModuleList(motion_modules)
        self.downsamplers = nn.ModuleList([Downsample3D(out_channels, use_conv=True, out_channels=out_channels,
                                                        padding=1, name="op")]) if add_downsample else None
        self.gradient_checkpointing = False

    def run(self, eng: Engine, x, si: StepInputs, motion_scale_reaches_audio: bool):
        outs = []
        for resnet, attn, audio, motion in zip(self.resnets, self.attentions, self.audio_modules, self.motion_modules):
            x = resnet.run(eng, x, None, si.temb_silu, si.frames)
            x = attn.run(eng, x, si.frames, clip_b=si.clip, seg2_index=si.seg2_index)
            if audio is not None:
                # fact 4: motion_scale only arrives through the gradient-checkpointing branch
                # (unet_3d_blocks.py:563-572 vs :591-600)
                scale = si.scale if motion_scale_reaches_audio else (1.0, 1.0, 1.0)
                x = audio.run(eng, x, si.frames, audio_rows=si.audio_rows, masks=si.masks[self.depth], scale=scale)
            if motion is not None:
                x = motion.run(eng, x, si.frames, si.shard)
            outs.append(x)
        if self.downsamplers is not None:
            x = self.downsamplers[0].run(eng, x)
            outs.append(x)
        return x, outs


class DownBlock3D(nn.Module):
    def __init__(self, in_channels, out_channels, temb_channels, num_layers=1, resnet_eps=1e-6, resnet_groups=32,
                 add_downsample=True, use_motion_module=None, motion_module_type=None, motion_module_kwargs=None):
        super().__init__()
        resnets, motion_modules = [], []
        for i in range(num_layers):
            in_c = in_channels if i == 0 else out_channels
            resnets.append(_resnet(in_c, out_channels, temb_channels, resnet_eps, resnet_groups))
            motion_modules.append(get_motion_module(out_channels, motion_module_type, motion_module_kwargs)
                                  if use_motion_module else None)
        self.resnets = nn.ModuleList(resnets)
        self.motion_modules = nn.ModuleList(motion_modules)
        self.downsamplers = nn.ModuleList([Downsample3D(out_channels, use_conv=True, out_channels=out_channels,
                                                        padding=1, name="op")]) if add_downsample else None
        self.gradient_checkpointing = False

    def run(self, eng: Engine, x, si: StepInputs):
        outs = []
        for resnet, motion in zip(self.resnets, self.motion_modules):
            x = resnet.run(eng, x, None, si.temb_silu, si.frames)
            if motion is not None:
                x = motion.run(eng, x, si.frames, si.shard)
            outs.append(x)
        if self.downsamplers is not None:
            x = self.downsamplers[0].run(eng, x)
            outs.append(x)
        return x, outs


class UNetMidBlock3DCrossAttn(nn.Module):
    """resnet -> spatial transformer -> motion -> resnet; no audio module here (unet_3d.py:176-196)."""

    def __init__(self, in_channels, temb_channels, num_layers=1, resnet_eps=1e-6, resnet_groups=32,
                 attn_num_head_channels=1, cross_attention_dim=1280, use_motion_module=None, motion_module_type=None,
                 motion_module_kwargs=None, name=None):
        super().__init__()
        self.has_cross_attention = True
        resnets = [_resnet(in_channels, in_channels, temb_channels, resnet_eps, resnet_groups)]
        attentions, motion_modules, audio_modules = [], [], []
        for i in range(num_layers):
            attentions.append(_spatial(attn_num_head_channels, in_channels, cross_attention_dim, resnet_groups,
                                       name=f"{name}_{i}_TransformerModel" if name else None))
            audio_modules.append(None)
            motion_modules.append(get_motion_module(in_channels, motion_module_type, motion_module_kwargs)
                                  if use_motion_module else None)
            resnets.append(_resnet(in_channels, in_channels, temb_channels, resnet_eps, resnet_groups))
        self.attentions = nn.ModuleList(attentions)
        self.resnets = nn.ModuleList(resnets)
        self.audio_modules = nn.ModuleList(audio_modules)
        self.motion_modules = nn.ModuleList(motion_modules)

    def run(self, eng: Engine, x, si: StepInputs):
        x = self.resnets[0].run(eng, x, None, si.temb_silu, si.frames)
        for attn, resnet, motion in zip(self.attentions, self.resnets[1:], self.motion_modules):
            x = attn.run(eng, x, si.frames, clip_b=si.clip, seg2_index=si.seg2_index)
            if motion is not None:
                x = motion.run(eng, x, si.frames, si.shard)
            x = resnet.run(eng, x, None, si.temb_silu, si.frames)
        return x


class CrossAttnUpBlock3D(nn.Module):
    def __init__(self, in_channels, out_channels, prev_output_channel, temb_channels, num_layers=1, resnet_eps=1e-6,
                 resnet_groups=32, attn_num_head_channels=1, cross_attention_dim=1280, add_upsample=True,
                 use_motion_module=None, motion_module_type=None, motion_module_kwargs=None, name=None):
        super().__init__()
        self.has_cross_attention = True
        resnets, attentions, audio_modules, motion_modules = [], [], [], []
        for i in range(num_layers):
            skip_c = in_channels if i == num_layers - 1 else out_channels
            in_c = prev_output_channel if i == 0 else out_channels
            resnets.append(_resnet(in_c + skip_c, out_channels, temb_channels, resnet_eps, resnet_groups))
            attentions.append(_spatial(attn_num_head_channels, out_channels, cross_attention_dim, resnet_groups,
                                       name=f"{name}_{i}_TransformerModel" if name else None))
            audio_modules.append(None)   # get_up_block is never handed use_audio_module (unet_3d.py:230-256)
            motion_modules.append(get_motion_module(out_channels, motion_module_type, motion_module_kwargs)
                                  if use_motion_module else None)
        self.attentions = nn.ModuleList(attentions)
        self.resnets = nn.ModuleList(resnets)
        self.audio_modules = nn.ModuleList(audio_modules)
        self.motion_modules = nn.ModuleList(motion_modules)
        self.upsamplers = nn.ModuleList([Upsample3D(out_channels, use_conv=True, out_channels=out_channels)]) \
            if add_upsample else None
        self.gradient_checkpointing = False

    def run(self, eng: Engine, x, skips: List[torch.Tensor], si: StepInputs):
        for resnet, attn, motion in zip(self.resnets, self.attentions, self.motion_modules):
            x = resnet.run(eng, x, skips.pop(), si.temb_silu, si.frames)
            x = attn.run(eng, x, si.frames, clip_b=si.clip, seg2_index=si.seg2_index)
            if motion is not None:
                x = motion.run(eng, x, si.frames, si.shard)
        if self.upsamplers is not None:
            x = self.upsamplers[0].run(eng, x)
        return x


class UpBlock3D(nn.Module):
    def __init__(self, in_channels, prev_output_channel, out_channels, temb_channels, num_layers=1, resnet_eps=1e-6,
                 resnet_groups=32, add_upsample=True, use_motion_module=None, motion_module_type=None,
                 motion_module_kwargs=None):
        super().__init__()
        resnets, motion_modules = [], []
        for i in range(num_layers):
            skip_c = in_channels if i == num_layers - 1 else out_channels
            in_c = prev_output_channel if i == 0 else out_channels
            resnets.append(_resnet(in_c + skip_c, out_channels, temb_channels, resnet_eps, resnet_groups))
            motion_modules.append(get_motion_module(out_channels, motion_module_type, motion_module_kwargs)
                                  if use_motion_module else None)
        self.resnets = nn.ModuleList(resnets)
        self.motion_modules = nn.ModuleList(motion_modules)
        self.upsamplers = nn.ModuleList([Upsample3D(out_channels, use_conv=True, out_channels=out_channels)]) \
            if add_upsample else None
        self.gradient_checkpointing = False

    def run(self, eng: Engine, x, skips: List[torch.Tensor], si: StepInputs):
        for resnet, motion in zip(self.resnets, self.motion_modules):
            x = resnet.run(eng, x, skips.pop(), si.temb_silu, si.frames)
            if motion is not None:
                x = motion.run(eng, x, si.frames, si.shard)
        if self.upsamplers is not None:
            x = self.upsamplers[0].run(eng, x)
        return x


def get_down_block(down_block_type, num_layers, in_channels, out_channels, temb_channels, add_downsample, resnet_eps,
                   resnet_act_fn, attn_num_head_channels, resnet_groups=None, cross_attention_dim=None,
                   audio_attention_dim=None, downsample_padding=None, use_motion_module=None, motion_module_type=None,
                   motion_module_kwargs=None, use_audio_module=None, depth=0, stack_enable_blocks_name=None,
                   stack_enable_blocks_depth=None, name_index=None, **unused):
    down_block_type = down_block_type[7:] if down_block_type.startswith("UNetRes") else down_block_type
    if down_block_type == "DownBlock3D":
        return DownBlock3D(in_channels, out_channels, temb_channels, num_layers=num_layers, resnet_eps=resnet_eps,
                           resnet_groups=resnet_groups, add_downsample=add_downsample, use_motion_module=use_motion_module,
                           motion_module_type=motion_module_type, motion_module_kwargs=motion_module_kwargs)
    if down_block_type == "CrossAttnDownBlock3D":
        if cross_attention_dim is None:
            raise ValueError("cross_attention_dim must be specified for CrossAttnDownBlock3D")
        return CrossAttnDownBlock3D(in_channels, out_channels, temb_channels, num_layers=num_layers, resnet_eps=resnet_eps,
                                    resnet_groups=resnet_groups, attn_num_head_channels=attn_num_head_channels,
                                    cross_attention_dim=cross_attention_dim, audio_attention_dim=audio_attention_dim,
                                    add_downsample=add_downsample, use_motion_module=use_motion_module,
                                    use_audio_module=use_audio_module, depth=depth,
                                    stack_enable_blocks_name=stack_enable_blocks_name,
                                    stack_enable_blocks_depth=stack_enable_blocks_depth,
                                    motion_module_type=motion_module_type, motion_module_kwargs=motion_module_kwargs,
                                    name=f"CrossAttnDownBlock_{name_index}_" if name_index is not None else None)
    raise ValueError(f"{down_block_type} does not exist.")


def get_up_block(up_block_type, num_layers, in_channels, out_channels, prev_output_channel, temb_channels, add_upsample,
                 resnet_eps, resnet_act_fn, attn_num_head_channels, resnet_groups=None, cross_attention_dim=None,
                 use_motion_module=None, motion_module_type=None, motion_module_kwargs=None, name_index=None, **unused):
    up_block_type = up_block_type[7:] if up_block_type.startswith("UNetRes") else up_block_type
    if up_block_type == "UpBlock3D":
        return UpBlock3D(in_channels, prev_output_channel, out_channels, temb_channels, num_layers=num_layers,
                         resnet_eps=resnet_eps, resnet_groups=resnet_groups, add_upsample=add_upsample,
                         use_motion_module=use_motion_module, motion_module_type=motion_module_type,
                         motion_module_kwargs=motion_module_kwargs)
    if up_block_type == "CrossAttnUpBlock3D":
        if cross_attention_dim is None:
            raise ValueError("cross_attention_dim must be specified for CrossAttnUpBlock3D")
        return CrossAttnUpBlock3D(in_channels, out_channels, prev_output_channel, temb_channels, num_layers=num_layers,
                                  resnet_eps=resnet_eps, resnet_groups=resnet_groups,
                                  attn_num_head_channels=attn_num_head_channels, cross_attention_dim=cross_attention_dim,
                                  add_upsample=add_upsample, use_motion_module=use_motion_module,
                                  motion_module_type=motion_module_type, motion_module_kwargs=motion_module_kwargs,
                                  name=f"CrossAttnUpBlock_{name_index}_" if name_index is not None else None)
    raise ValueError(f"{up_block_type} does not exist.")
