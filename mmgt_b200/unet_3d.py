"""Host mirror of src/models/unet_3d.py: ``UNet3DConditionModel`` with the reference's constructor
arguments, state-dict keys / shapes and ``forward`` signature, executing on the sm_100a kernels.

Differences that are deliberate (and invisible through the reference API):
  * activations live channels-last as (N = B*F, H, W, C); NCFHW only exists at the boundary;
  * the reference-feature (bank) K/V projections are cached per video, the CFG "uncond re-do" is a
    per-frame key-length switch inside one attention launch, CLIP cross-attention (1 token) is a vector add;
  * compute dtype is float32 or bfloat16 (the reference's fp16 maps to bf16); parameters may stay float32
    while kernels run in bf16 (``set_compute_dtype``).
"""
import json
import os
from dataclasses import dataclass
from pathlib import Path
from typing import List, Optional, Sequence, Tuple, Union

import torch
import torch.nn as nn

from .kernels import Engine, get_engine
from .packing import Pack, f32
from .resnet import InflatedConv3d, InflatedGroupNorm
from .unet_3d_blocks import (CrossAttnDownBlock3D, StepInputs, UNetMidBlock3DCrossAttn, get_down_block, get_up_block)


class FrozenConfig(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


@dataclass
class UNet3DConditionOutput:
    sample: torch.Tensor

    def __getitem__(self, i):
        return (self.sample,)[i]


class Timesteps(nn.Module):
    """Parameter-free sinusoidal projection (diffusers Timesteps); evaluated by mmgt_timestep_embedding."""

    def __init__(self, num_channels, flip_sin_to_cos, downscale_freq_shift):
        super().__init__()
        self.num_channels, self.flip_sin_to_cos, self.downscale_freq_shift = num_channels, flip_sin_to_cos, downscale_freq_shift


class TimestepEmbedding(nn.Module):
    def __init__(self, in_channels, time_embed_dim):
        super().__init__()
        self.linear_1 = nn.Linear(in_channels, time_embed_dim)
        self.act = nn.SiLU()
        self.linear_2 = nn.Linear(time_embed_dim, time_embed_dim)
        self._pack = Pack()

    def run(self, eng: Engine, t_emb):
        w1, b1, w2, b2 = self._pack.get(
            eng, [self.linear_1.weight, self.linear_1.bias, self.linear_2.weight, self.linear_2.bias],
            lambda: (f32(self.linear_1.weight, eng), f32(self.linear_1.bias, eng), f32(self.linear_2.weight, eng),
                     f32(self.linear_2.bias, eng)))
        h = eng.silu_f32(eng.gemm(t_emb, w1, bias=b1, dtype=torch.float32))
        return eng.gemm(h, w2, bias=b2, dtype=torch.float32)


class TimeProjections:
    """silu(time embedding) pushed through the ``time_emb_proj`` of ALL resnets by one GEMV: (B, sum of widths) float32.
    A resnet's (B, Cout) rowbias is a column slice (``of``); ``rows`` selects batch rows for single-branch forwards.
    Computed once per DDIM step (the timestep is the same for every context window, pipeline_pose2vid_long.py:554-620)
    instead of 2 + 22 GEMVs per window forward."""

    def __init__(self, all_proj: torch.Tensor, offsets: dict):
        self.all_proj, self.offsets = all_proj, offsets

    def of(self, resnet) -> torch.Tensor:
        o, n = self.offsets[id(resnet)]
        return self.all_proj[:, o:o + n]

    def rows(self, B: int) -> "TimeProjections":
        return self if B == self.all_proj.shape[0] else TimeProjections(self.all_proj[:B], self.offsets)


_INIT_DEFAULTS = dict(
    sample_size=None, in_channels=4, out_channels=4, flip_sin_to_cos=True, freq_shift=0,
    down_block_types=("CrossAttnDownBlock3D", "CrossAttnDownBlock3D", "CrossAttnDownBlock3D", "DownBlock3D"),
    mid_block_type="UNetMidBlock3DCrossAttn",
    up_block_types=("UpBlock3D", "CrossAttnUpBlock3D", "CrossAttnUpBlock3D", "CrossAttnUpBlock3D"),
    only_cross_attention=False, block_out_channels=(320, 640, 1280, 1280), layers_per_block=2, downsample_padding=1,
    mid_block_scale_factor=1, act_fn="silu", norm_num_groups=32, norm_eps=1e-5, cross_attention_dim=1280,
    attention_head_dim=8, dual_cross_attention=False, use_linear_projection=False, class_embed_type=None,
    num_class_embeds=None, upcast_attention=False, resnet_time_scale_shift="default", use_inflated_groupnorm=False,
    task_type="action", mode=None, use_motion_module=False, motion_module_resolutions=(1, 2, 4, 8),
    motion_module_mid_block=False, motion_module_decoder_only=False, motion_module_type=None, motion_module_kwargs={},
    unet_use_cross_frame_attention=None, unet_use_temporal_attention=None, use_audio_module=False,
    audio_attention_dim=768, stack_enable_blocks_name=None, stack_enable_blocks_depth=None)


class UNet3DConditionModel(nn.Module):
    """Same constructor keywords as the reference (unet_3d.py:37-90)."""

    _supports_gradient_checkpointing = True
    config_name = "config.json"

    def __init__(self, **kwargs):
        super().__init__()
        unknown = set(kwargs) - set(_INIT_DEFAULTS)
        if unknown:
            raise TypeError(f"unexpected arguments {sorted(unknown)}")
        cfg = dict(_INIT_DEFAULTS)
        cfg.update(kwargs)
        self._internal_dict = FrozenConfig(cfg)
        c = self._internal_dict
        if c.dual_cross_attention or c.use_linear_projection or c.class_embed_type is not None \
                or c.num_class_embeds is not None or c.resnet_time_scale_shift != "default" or c.mid_block_scale_factor != 1:
            raise NotImplementedError("configuration outside the SD-1.5 / animation.yaml hot path")
        if not c.use_inflated_groupnorm:
            raise NotImplementedError("use_inflated_groupnorm=False (animation.yaml:48 sets it true)")
        boc = list(c.block_out_channels)
        self.sample_size = c.sample_size
        time_embed_dim = boc[0] * 4
        self.conv_in = InflatedConv3d(c.in_channels, boc[0], kernel_size=3, padding=(1, 1))
        self.time_proj = Timesteps(boc[0], c.flip_sin_to_cos, c.freq_shift)
        self.time_embedding = TimestepEmbedding(boc[0], time_embed_dim)
        self.class_embedding = None
        self.down_blocks = nn.ModuleList([])
        self.mid_block = None          # plain attribute for now: keeps the reference's module order down->up->mid
        self.up_blocks = nn.ModuleList([])
        heads = c.attention_head_dim
        heads = (heads,) * len(c.down_block_types) if isinstance(heads, int) else tuple(heads)

        out_c = boc[0]
        for i, btype in enumerate(c.down_block_types):
            in_c, out_c = out_c, boc[i]
            res = 2 ** i
            self.down_blocks.append(get_down_block(
                btype, num_layers=c.layers_per_block, in_channels=in_c, out_channels=out_c, temb_channels=time_embed_dim,
                add_downsample=i != len(boc) - 1, resnet_eps=c.norm_eps, resnet_act_fn=c.act_fn,
                resnet_groups=c.norm_num_groups, cross_attention_dim=c.cross_attention_dim,
                attn_num_head_channels=heads[i], downsample_padding=c.downsample_padding,
                use_motion_module=c.use_motion_module and (res in c.motion_module_resolutions)
                and (not c.motion_module_decoder_only),
                motion_module_type=c.motion_module_type, motion_module_kwargs=c.motion_module_kwargs,
                use_audio_module=c.use_audio_module, audio_attention_dim=c.audio_attention_dim, depth=i,
                stack_enable_blocks_name=c.stack_enable_blocks_name, stack_enable_blocks_depth=c.stack_enable_blocks_depth,
                name_index=None if c.task_type == "action" else i))

        if c.mid_block_type != "UNetMidBlock3DCrossAttn":
            raise ValueError(f"unknown mid_block_type : {c.mid_block_type}")
        self.mid_block = UNetMidBlock3DCrossAttn(
            in_channels=boc[-1], temb_channels=time_embed_dim, resnet_eps=c.norm_eps, resnet_groups=c.norm_num_groups,
            attn_num_head_channels=heads[-1], cross_attention_dim=c.cross_attention_dim,
            use_motion_module=c.use_motion_module and c.motion_module_mid_block, motion_module_type=c.motion_module_type,
            motion_module_kwargs=c.motion_module_kwargs, name=None if c.task_type == "action" else "MidBlock")

        self.num_upsamplers = 0
        rboc, rheads = list(reversed(boc)), list(reversed(heads))
        out_c = rboc[0]
        for i, btype in enumerate(c.up_block_types):
            res = 2 ** (3 - i)
            prev_c, out_c = out_c, rboc[i]
            in_c = rboc[min(i + 1, len(boc) - 1)]
            add_up = i != len(boc) - 1
            self.num_upsamplers += int(add_up)
            self.up_blocks.append(get_up_block(
                btype, num_layers=c.layers_per_block + 1, in_channels=in_c, out_channels=out_c, prev_output_channel=prev_c,
                temb_channels=time_embed_dim, add_upsample=add_up, resnet_eps=c.norm_eps, resnet_act_fn=c.act_fn,
                resnet_groups=c.norm_num_groups, cross_attention_dim=c.cross_attention_dim,
                attn_num_head_channels=rheads[i],
                use_motion_module=c.use_motion_module and (res in c.motion_module_resolutions),
                motion_module_type=c.motion_module_type, motion_module_kwargs=c.motion_module_kwargs,
                name_index=None if c.task_type == "action" else i))

        self.conv_norm_out = InflatedGroupNorm(num_channels=boc[0], num_groups=c.norm_num_groups, eps=c.norm_eps)
        self.conv_act = nn.SiLU()
        self.conv_out = InflatedConv3d(boc[0], c.out_channels, kernel_size=3, padding=1)
        self.mode = c.mode
        self.compute_dtype: Optional[torch.dtype] = None   # None => follow the parameter dtype
        # None => the reference's accidental rule (fact 4): motion_scale reaches MM-HAA only when
        # `self.training and gradient_checkpointing`; True/False force it.
        self.apply_motion_scale: Optional[bool] = None
        self._idx_cache = {}
        self._time_pack = Pack()

    # ------------------------------------------------------------------ diffusers-style plumbing
    @property
    def config(self):
        return self._internal_dict

    def __getattr__(self, name):
        try:
            return super().__getattr__(name)
        except AttributeError:
            d = self.__dict__.get("_internal_dict")
            if d is not None and name in d:
                return d[name]
            raise

    @property
    def device(self):
        return next(self.parameters()).device

    @property
    def dtype(self):
        return next(self.parameters()).dtype

    @classmethod
    def from_config(cls, config, **kwargs):
        cfg = {k: v for k, v in dict(config).items() if not k.startswith("_")}
        cfg.update(kwargs)
        model = cls(**{k: v for k, v in cfg.items() if k in _INIT_DEFAULTS})
        extra = {k: v for k, v in cfg.items() if k not in _INIT_DEFAULTS}
        extra.setdefault("center_input_sample", False)
        model._internal_dict = FrozenConfig({**model._internal_dict, **extra})
        return model

    @classmethod
    def load_config(cls, path):
        with open(path) as f:
            return json.load(f)

    def _set_gradient_checkpointing(self, module, value=False):
        if hasattr(module, "gradient_checkpointing"):
            module.gradient_checkpointing = value

    def enable_gradient_checkpointing(self):
        for m in self.modules():
            self._set_gradient_checkpointing(m, True)

    def disable_gradient_checkpointing(self):
        for m in self.modules():
            self._set_gradient_checkpointing(m, False)

    @property
    def attn_processors(self):
        return {}

    def set_attn_processor(self, processor):
        return None

    def set_attention_slice(self, slice_size):
        return None

    def set_compute_dtype(self, dtype: Optional[torch.dtype]):
        """Run the kernels in ``dtype`` (float32 / bfloat16) regardless of the parameter dtype."""
        self.compute_dtype = dtype
        return self

    @classmethod
    def from_pretrained_2d(cls, pretrained_model_path, motion_module_path, subfolder=None, unet_additional_kwargs=None,
                           mm_zero_proj_out=False):
        """Same contract as unet_3d.py:627-718: SD-1.5 2-D weights + motion-module checkpoint, strict=False."""
        pretrained_model_path = Path(pretrained_model_path)
        motion_module_path = Path(motion_module_path)
        if subfolder is not None:
            pretrained_model_path = pretrained_model_path.joinpath(subfolder)
        config_file = pretrained_model_path / "config.json"
        if not (config_file.exists() and config_file.is_file()):
            raise RuntimeError(f"{config_file} does not exist or is not a file")
        unet_config = cls.load_config(config_file)
        unet_config["down_block_types"] = ["CrossAttnDownBlock3D"] * 3 + ["DownBlock3D"]
        unet_config["up_block_types"] = ["UpBlock3D"] + ["CrossAttnUpBlock3D"] * 3
        unet_config["mid_block_type"] = "UNetMidBlock3DCrossAttn"
        model = cls.from_config(unet_config, **(unet_additional_kwargs or {}))
        st_file = pretrained_model_path / "diffusion_pytorch_model.safetensors"
        bin_file = pretrained_model_path / "diffusion_pytorch_model.bin"
        if st_file.exists():
            from safetensors.torch import load_file
            state_dict = load_file(str(st_file), device="cpu")
        elif bin_file.exists():
            state_dict = torch.load(bin_file, map_location="cpu", weights_only=True)
        else:
            raise FileNotFoundError(f"no weights file found in {pretrained_model_path}")
        if motion_module_path.exists() and motion_module_path.is_file():
            if motion_module_path.suffix.lower() in (".pth", ".pt", ".ckpt"):
                motion_sd = torch.load(motion_module_path, map_location="cpu", weights_only=True)
            elif motion_module_path.suffix.lower() == ".safetensors":
                from safetensors.torch import load_file
                motion_sd = load_file(str(motion_module_path), device="cpu")
            else:
                raise RuntimeError(f"unknown file format for motion module weights: {motion_module_path.suffix}")
            if mm_zero_proj_out:
                motion_sd = {k: v for k, v in motion_sd.items() if "proj_out" not in k}
            state_dict.update(motion_sd)
        model.load_state_dict(state_dict, strict=False)
        return model

    # ------------------------------------------------------------------ helpers
    def _engine(self, device) -> Engine:
        dt = self.compute_dtype or self.dtype
        if dt == torch.float16:
            dt = torch.bfloat16
        return get_engine(device, dt)

    def time_projections(self, eng: Engine, timestep, B: int) -> TimeProjections:
        """unet_3d.py:481-502 + resnet.py:226 for every resnet at once; float32 end to end.  ``timestep``: python number,
        0-d / (1,) / (B,) tensor (a device tensor keeps the call CUDA-graph capturable)."""
        from .resnet import ResnetBlock3D
        dev = eng.device
        if torch.is_tensor(timestep):
            t = timestep.to(device=dev, dtype=torch.float32).reshape(-1)
        else:
            t = torch.tensor([float(timestep)], device=dev, dtype=torch.float32)
        if t.numel() == 1:
            t = t.expand(B).contiguous()
        tp = self.time_proj
        emb = self.time_embedding.run(eng, eng.timestep_embedding(t, tp.num_channels, tp.flip_sin_to_cos,
                                                                  tp.downscale_freq_shift))
        temb_silu = eng.silu_f32(emb)
        resnets = [m for m in self.modules() if isinstance(m, ResnetBlock3D)]
        params = [p for r in resnets for p in (r.time_emb_proj.weight, r.time_emb_proj.bias)]

        def build():
            w = torch.cat([f32(r.time_emb_proj.weight, eng) for r in resnets], dim=0).contiguous()
            b = torch.cat([f32(r.time_emb_proj.bias, eng) for r in resnets], dim=0).contiguous()
            offs, o = {}, 0
            for r in resnets:
                n = r.time_emb_proj.weight.shape[0]
                offs[id(r)] = (o, n)
                o += n
            return w, b, offs
        w, b, offs = self._time_pack.get(eng, params, build)
        return TimeProjections(eng.gemm(temb_silu, w, bias=b, dtype=torch.float32), offs)

    def spatial_blocks(self):
        """The 16 spatial transformer blocks in module order (what torch_dfs + isinstance finds)."""
        from .attention import TemporalBasicTransformerBlock
        return [m for m in self.modules() if isinstance(m, TemporalBasicTransformerBlock)]

    def _seg2_index(self, eng, B, F, ref_index):
        key = (str(eng.device), B, F, tuple(-1 if r is None else int(r) for r in ref_index))
        if key not in self._idx_cache:
            idx = torch.tensor([v for v in key[3] for _ in range(F)], dtype=torch.int32, device=eng.device)
            idx._n_seg2 = sum(F for v in key[3] if v >= 0)     # host-side count for the FLOP accounting of bench.py
            self._idx_cache[key] = idx
        return self._idx_cache[key]

    def _motion_scale_reaches_audio(self) -> bool:
        if self.apply_motion_scale is not None:
            return bool(self.apply_motion_scale)
        blk = next((b for b in self.down_blocks if isinstance(b, CrossAttnDownBlock3D)), None)
        return bool(self.training and blk is not None and blk.gradient_checkpointing)

    # ------------------------------------------------------------------ forward
    @torch.no_grad()
    def forward(self, sample, timestep, encoder_hidden_states, audio_embedding=None, class_labels=None,
                mask_cond_fea=None, pose_cond_fea=None, attention_mask=None, full_mask=None, face_mask=None,
                body_mask=None, motion_scale=None, down_block_additional_residuals=None,
                mid_block_additional_residual=None, return_dict: bool = True, ref_index: Optional[Sequence] = None):
        """Reference signature (unet_3d.py:425-443) plus ``ref_index``: per batch sample, the row of the
        reference bank its frames attend to, or None for plain self-attention.  Default = what
        ReferenceAttentionControl configured (CFG: first half None, second half its own row)."""
        if attention_mask is not None or down_block_additional_residuals is not None \
                or mid_block_additional_residual is not None or class_labels is not None:
            raise NotImplementedError("attention_mask / additional residuals / class labels are not on the hot path")
        if sample.dim() != 5:
            raise ValueError(f"Expected sample to have ndim=5, but got ndim={sample.dim()}")
        B, Cin, F, H, W = sample.shape
        eng = self._engine(sample.device)
        if self.config.get("center_input_sample", False):
            sample = 2 * sample - 1.0
        x = eng.ncfhw_to_tokens(sample)
        pose = eng.ncfhw_to_tokens(pose_cond_fea) if pose_cond_fea is not None else None
        out = self.forward_tokens(eng, x, timestep, encoder_hidden_states, audio_embedding, pose, full_mask, face_mask,
                                  body_mask, motion_scale, B, F, ref_index)
        res = eng.tokens_to_ncfhw(out, B, F, sample.dtype if sample.dtype in (torch.float32, torch.bfloat16) else torch.float32)
        if res.dtype != sample.dtype:
            res = res.to(sample.dtype)
        if not return_dict:
            return (res,)
        return UNet3DConditionOutput(sample=res)

    def forward_tokens(self, eng: Engine, x, timestep, encoder_hidden_states, audio_embedding, pose, full_mask, face_mask,
                       body_mask, motion_scale, B: int, F: int, ref_index=None, shard=None, time_proj=None):
        """Channels-last core: x (N,H,W,4), pose (N,H,W,320) or None -> (N,H,W,4) in the run dtype.
        ``shard`` (frame_shard.FrameShardGroup): x / pose / audio / masks hold only this rank's F frames of a window
        whose k*F frames are spread over k ranks; the motion modules exchange rows with the peers.
        ``time_proj``: TimeProjections of this timestep computed by the caller (once per DDIM step)."""
        si = self._step_inputs(eng, x, timestep, encoder_hidden_states, audio_embedding, full_mask, face_mask, body_mask,
                               motion_scale, B, F, ref_index, shard, time_proj)
        reaches = self._motion_scale_reaches_audio()
        x, skips = self._run_down(eng, x, pose, si, reaches, 0, len(self.down_blocks))
        x = self.mid_block.run(eng, x, si)
        return self._run_up(eng, x, skips, si, 0, len(self.up_blocks), final=True)

    def forward_tokens_group(self, eng: Engine, units, timestep, motion_scale, time_proj=None, deep_from: int = 2,
                             level_batch=None):
        """Several independent forwards (context windows / CFG branches of one DDIM step) with the levels batched differently.
        ``units``: list of dicts(x, pose, ehs, audio, full, face, body, B, F, ref) -- the arguments of ``forward_tokens``.
        ``level_batch[l]`` = how many consecutive units run down block l (and the matching up block) as ONE batch; the values
        must not decrease with depth.  Default (``deep_from`` = 2): ``[1, 1, all, all]`` -- the 64x64 and 32x32 levels at
        512x512 run unit by unit, so their tensors stay L2-sized, while the 16x16 / 8x8 levels, the mid block and the first up
        blocks run ONCE on the frames of all units concatenated (a 12-frame window gives those GEMMs M = 6144 / 1536 rows,
        48 / 12 tiles for 148 SMs; ten windows fill the machine and amortise ~1200 launches).  Consecutive levels with the
        same batch form a stage; a stage runs its down blocks chunk by chunk, hands the concatenated result to the next
        stage and runs its up blocks chunk by chunk on the rows that come back.  Every layer is per sample / per frame /
        per (sample, pixel), so the result is the same as ``forward_tokens`` unit by unit.  -> list of outputs."""
        nblk = len(self.down_blocks)
        if level_batch is None:
            level_batch = [1] * min(max(int(deep_from), 0), nblk) + [len(units)] * max(nblk - max(int(deep_from), 0), 0)
        level_batch = [max(1, min(int(g), len(units))) for g in level_batch]
        if len(level_batch) != nblk or any(a > b for a, b in zip(level_batch, level_batch[1:])):
            raise ValueError(f"level_batch needs {nblk} non-decreasing entries, got {level_batch}")
        if len(units) == 1 or level_batch[-1] == 1:
            return [self.forward_tokens(eng, u["x"], timestep, u["ehs"], u["audio"], u["pose"], u["full"], u["face"], u["body"],
                                        motion_scale, u["B"], u["F"], ref_index=u["ref"], time_proj=time_proj) for u in units]
        F = units[0]["F"]
        if any(u["F"] != F for u in units):
            raise ValueError("forward_tokens_group: all units must have the same number of frames per sample")
        reaches = self._motion_scale_reaches_audio()
        stages = []                                    # (first down block, one past the last, units per batch)
        for lvl, g in enumerate(level_batch):
            if stages and stages[-1][2] == g:
                stages[-1] = (stages[-1][0], lvl + 1, g)
            else:
                stages.append((lvl, lvl + 1, g))
        sis = [self._step_inputs(eng, u["x"], timestep, u["ehs"], u["audio"], u["full"], u["face"], u["body"], motion_scale,
                                 u["B"], u["F"], u["ref"], None, time_proj) for u in units]
        nrows = [u["B"] * u["F"] for u in units]       # frames (= leading rows of every activation) per unit
        have_audio = sis[0].audio_rows is not None

        def merged(ids, lo):
            """StepInputs of the units ``ids`` run as one batch from down block ``lo`` on."""
            if len(ids) == 1:
                return sis[ids[0]]
            part = [sis[i] for i in ids]
            masks = None
            if have_audio:
                masks = [None if lvl < lo else tuple(torch.cat([si.masks[lvl][r] for si in part]) for r in range(3))
                         for lvl in range(len(part[0].masks))]
            tp0 = part[0].temb_silu
            Bt = sum(units[i]["B"] for i in ids)
            tproj = type(tp0)(tp0.all_proj[:1].expand(Bt, -1).contiguous(), tp0.offsets)   # same timestep for every sample
            seg2 = None
            if part[0].seg2_index is not None:
                seg2 = torch.cat([si.seg2_index for si in part])
                seg2._n_seg2 = sum(getattr(si.seg2_index, "_n_seg2", si.seg2_index.numel()) for si in part)   # bench FLOPs
            return StepInputs(frames=F, temb_silu=tproj, clip=torch.cat([si.clip for si in part]), seg2_index=seg2,
                              audio_rows=torch.cat([si.audio_rows for si in part]) if have_audio else None, masks=masks,
                              scale=part[0].scale, shard=None)

        def run_stage(s, ids, X):
            """Down blocks of stage ``s`` and deeper, mid block, up blocks back to stage ``s`` for the units ``ids``;
            ``X``: their concatenated input rows (None for the first stage: conv_in reads the units' own tensors)."""
            lo, hi, g = stages[s]
            last = s == len(stages) - 1
            state, n0 = [], 0
            for c0 in range(0, len(ids), g):
                part = ids[c0:c0 + g]
                n = sum(nrows[i] for i in part)
                si = merged(part, lo)
                if lo == 0:
                    xp = units[part[0]]["x"] if len(part) == 1 else torch.cat([units[i]["x"] for i in part])
                    pose = units[part[0]]["pose"] if len(part) == 1 else torch.cat([units[i]["pose"] for i in part])
                    x, skips = self._run_down(eng, xp, pose, si, reaches, 0, hi)
                else:
                    xp = X[n0:n0 + n]
                    x, skips = self._run_down(eng, xp, None, si, reaches, lo, hi, skips=[xp])
                if last:
                    x = self.mid_block.run(eng, x, si)
                state.append((si, x, skips, n))
                n0 += n
            if not last:
                Y = run_stage(s + 1, ids, state[0][1] if len(state) == 1 else torch.cat([st[1] for st in state]))
            outs, n0 = [], 0
            for si, x, skips, n in state:
                if not last:
                    skips.pop()                  # the last skip is the next stage's input itself (consumed inside its batch)
                    x = Y[n0:n0 + n]
                outs.append(self._run_up(eng, x, skips, si, nblk - hi, nblk - lo, final=(lo == 0)))
                assert len(skips) == 0
                n0 += n
            return outs if lo == 0 else (outs[0] if len(outs) == 1 else torch.cat(outs))

        outs = run_stage(0, list(range(len(units))), None)
        if stages[0][2] == 1:
            return outs
        res = []                                        # first-stage chunks hold several units: back to one output per unit
        it = iter(outs)
        for c0 in range(0, len(units), stages[0][2]):
            y, n0 = next(it), 0
            for i in range(c0, min(c0 + stages[0][2], len(units))):
                res.append(y[n0:n0 + nrows[i]])
                n0 += nrows[i]
        return res

    def _run_down(self, eng, x, pose, si, reaches, lo, hi, skips=None):
        if lo == 0:
            # pre-process: conv_in (+ pose features fused as the residual) (unet_3d.py:517-519)
            x = self.conv_in.run(eng, x, residual=pose)
            skips = [x]
        for blk in list(self.down_blocks)[lo:hi]:
            if isinstance(blk, CrossAttnDownBlock3D):
                x, outs = blk.run(eng, x, si, reaches)
            else:
                x, outs = blk.run(eng, x, si)
            skips += outs
        return x, skips

    def _run_up(self, eng, x, skips, si, lo, hi, final):
        for blk in list(self.up_blocks)[lo:hi]:
            x = blk.run(eng, x, skips, si)
        if not final:
            return x
        x = self.conv_norm_out.run(eng, x, None, silu=True)
        return self.conv_out.run(eng, x)

    def _step_inputs(self, eng: Engine, x, timestep, encoder_hidden_states, audio_embedding, full_mask, face_mask, body_mask,
                     motion_scale, B: int, F: int, ref_index=None, shard=None, time_proj=None):
        """Validation + the per-forward conditioning every block shares (StepInputs)."""
        N, H, W, _ = x.shape
        dev = eng.device
        down = 2 ** self.num_upsamplers
        if H % down or W % down:
            raise ValueError(f"latent size {H}x{W} must be divisible by {down} (the reference's upsample_size path for "
                             "other sizes, unet_3d.py:458-475, is not implemented)")
        # --- time embedding (unet_3d.py:481-502) and every resnet's time_emb_proj, float32 end to end
        temb_silu = time_proj.rows(B) if time_proj is not None else self.time_projections(eng, timestep, B)
        # --- conditioning
        if ref_index is None:
            ctl = getattr(self, "_reference_control", None)
            if ctl is not None and ctl.get("do_classifier_free_guidance", False):
                ref_index = [None] * (B // 2) + list(range(B // 2, B))
            else:
                ref_index = list(range(B))
        have_bank = any(len(b.bank) > 0 for b in self.spatial_blocks())
        if have_bank:
            nb = next(b.bank[0].shape[0] for b in self.spatial_blocks() if b.bank)
            ref_index = [None if r is None else min(int(r), nb - 1) for r in ref_index]
        seg2 = self._seg2_index(eng, B, F, ref_index) if have_bank else None
        clip = encoder_hidden_states.to(dev)
        if clip.shape[0] != B:
            raise ValueError("encoder_hidden_states batch must match sample batch")
        audio_rows, masks = None, None
        if self.config.use_audio_module:
            if audio_embedding is None or full_mask is None or face_mask is None or body_mask is None:
                raise ValueError("use_audio_module=True needs audio_embedding and the full/face/body masks")
            audio_rows = audio_embedding.to(device=dev, dtype=eng.dtype).reshape(N * audio_embedding.shape[2], -1).contiguous()
            masks = []
            for lvl in range(len(full_mask)):
                masks.append(tuple(m[lvl].to(device=dev, dtype=torch.float32).reshape(-1).contiguous()
                                   for m in (full_mask, face_mask, body_mask)))
        scale = tuple(float(s) for s in motion_scale) if motion_scale is not None else (1.0, 1.0, 1.0)
        si = StepInputs(frames=F, temb_silu=temb_silu, clip=clip, seg2_index=seg2, audio_rows=audio_rows, masks=masks,
                        scale=scale, shard=shard)
        return si
