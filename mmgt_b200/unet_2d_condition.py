"""ReferenceNet on the sm_100a kernels: host mirror of src/models/unet_2d_condition.py as the
hot path uses it -- ONE forward of the SD-1.5 2-D UNet on the reference-image latents at t = 0 in write mode
(pipeline_pose2vid_long.py:510-520), whose only product is the 16 reference-feature banks.

A 2-D UNet is the 3-D one of this package with one frame per sample and neither motion nor audio modules: ``InflatedConv3d``,
``InflatedGroupNorm``, ``Transformer3DModel`` and the spatial transformer block degenerate to their 2-D counterparts, and the
state-dict keys are the SD-1.5 UNet's (``down_blocks.i.resnets.j...``, ``attentions.j.transformer_blocks.0...``,
``downsamplers.0.conv``, ``upsamplers.0.conv``; no ``motion_modules`` / ``audio_modules``), so
``stable-diffusion-v1-5/unet`` weights load unchanged.  ``ReferenceAttentionControl(mode="write")`` flags the spatial blocks,
which then append ``norm1(hidden_states)`` to their bank (mutual_self_attention.py:139-148) and attend to themselves only.
"""
import json
from pathlib import Path

import torch

from .unet_3d import UNet3DConditionModel, UNet3DConditionOutput

_TO_3D = {"CrossAttnDownBlock2D": "CrossAttnDownBlock3D", "DownBlock2D": "DownBlock3D", "UpBlock2D": "UpBlock3D",
          "CrossAttnUpBlock2D": "CrossAttnUpBlock3D", "UNetMidBlock2DCrossAttn": "UNetMidBlock3DCrossAttn"}


class UNet2DConditionModel(UNet3DConditionModel):
    def __init__(self, **kwargs):
        kwargs = dict(kwargs)
        for key in ("down_block_types", "up_block_types"):
            if key in kwargs:
                kwargs[key] = tuple(_TO_3D.get(t, t) for t in kwargs[key])
        if "mid_block_type" in kwargs:
            kwargs["mid_block_type"] = _TO_3D.get(kwargs["mid_block_type"], kwargs["mid_block_type"])
        kwargs.update(use_inflated_groupnorm=True, use_motion_module=False, use_audio_module=False,
                      unet_use_cross_frame_attention=False, unet_use_temporal_attention=False)
        super().__init__(**kwargs)

    @classmethod
    def from_pretrained(cls, pretrained_model_path, subfolder=None, **unused):
        """SD-1.5 ``unet/config.json`` + ``diffusion_pytorch_model.{safetensors,bin}`` (scripts/pose2vid.py:146-149)."""
        path = Path(pretrained_model_path)
        if subfolder is not None:
            path = path / subfolder
        with open(path / "config.json") as f:
            model = cls.from_config(json.load(f))
        st, bn = path / "diffusion_pytorch_model.safetensors", path / "diffusion_pytorch_model.bin"
        if st.exists():
            from safetensors.torch import load_file
            sd = load_file(str(st), device="cpu")
        elif bn.exists():
            sd = torch.load(bn, map_location="cpu", weights_only=True)
        else:
            raise FileNotFoundError(f"no weights file found in {path}")
        model.load_state_dict(sd, strict=True)
        return model

    @torch.no_grad()
    def forward(self, sample, timestep, encoder_hidden_states, class_labels=None, timestep_cond=None, attention_mask=None,
                cross_attention_kwargs=None, added_cond_kwargs=None, down_block_additional_residuals=None,
                mid_block_additional_residual=None, encoder_attention_mask=None, return_dict: bool = True, **unused):
        """Reference signature (unet_2d_condition.py:872-890); sample (B, 4, h, w)."""
        if any(a is not None for a in (class_labels, timestep_cond, attention_mask, cross_attention_kwargs, added_cond_kwargs,
                                       down_block_additional_residuals, mid_block_additional_residual,
                                       encoder_attention_mask)):
            raise NotImplementedError("only (sample, timestep, encoder_hidden_states) are on the ReferenceNet path")
        if sample.dim() != 4:
            raise ValueError(f"Expected sample to have ndim=4, but got ndim={sample.dim()}")
        out = UNet3DConditionModel.forward(self, sample.unsqueeze(2), timestep, encoder_hidden_states, return_dict=False,
                                           ref_index=list(range(sample.shape[0])))[0].squeeze(2)
        return UNet3DConditionOutput(sample=out) if return_dict else (out,)
