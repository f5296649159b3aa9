"""Golden fixture for the ReferenceNet write pass.  TEST INFRASTRUCTURE; runs only where
/root/reference exists.

The reference's 2-D ``UNet2DConditionModel`` (src/models/unet_2d_condition.py) is built from diffusers' ``unet_2d_blocks``
primitives (``ResnetBlock2D``, ``Downsample2D`` ..., unet_2d_blocks.py:7-12), which are neither under /root/reference nor
installable here.  What pins the write pass instead is the reference's OWN code for everything that matters on this path:
its ``UNet3DConditionModel`` (src/models/unet_3d.py) built WITHOUT motion and audio modules and called on one frame per
sample -- per-frame convolutions, GroupNorms and spatial transformer blocks, i.e. the same network as the SD-1.5 2-D UNet with
the same state-dict keys -- under its own ``ReferenceAttentionControl(mode="write")`` (src/models/mutual_self_attention.py:
139-148: ``bank.append(norm1(hidden_states))``), fed like pipeline_pose2vid_long.py:510-520 feeds the ReferenceNet (latents
repeated for CFG, t = 0, [zeros; clip] as encoder_hidden_states).  Stored in tests/golden/refnet_tiny.npz: the 16 banks in
pairing order (stable sort by descending width) and the network output; inputs are regenerated from seeds.

Usage:  python -m oracle.make_golden_refnet
"""
import json
import os

import numpy as np
import torch

from . import reference_loader as RL
from .weights import make_state_dict

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
TINY = [64, 128, 256, 256]


def refnet_inputs(latent=16, seed=21):
    g = torch.Generator().manual_seed(seed)
    ref_latents = torch.randn(1, 4, latent, latent, generator=g)
    clip = torch.randn(1, 1, 768, generator=g)
    return ref_latents.repeat(2, 1, 1, 1), torch.cat([torch.zeros_like(clip), clip])


def main():
    mods = RL.load_reference_modules()
    cfg = dict(RL.SD15_CFG)
    cfg["block_out_channels"] = TINY
    extra = dict(RL.UNET_ADDITIONAL_KWARGS)
    extra.update(use_motion_module=False, use_audio_module=False)
    unet = mods["unet_3d"].UNet3DConditionModel.from_config(cfg, **extra)
    spec = [(k, tuple(v.shape)) for k, v in unet.state_dict().items()]
    sd = make_state_dict(spec, seed=8)
    unet.load_state_dict(sd, strict=True)
    unet.eval()
    msa, attn_mod = mods["mutual_self_attention"], mods["attention"]
    writer = msa.ReferenceAttentionControl(unet, do_classifier_free_guidance=True, mode="write", batch_size=1,
                                           fusion_blocks="full")
    x, ehs = refnet_inputs()
    with torch.no_grad():
        out = unet(x.unsqueeze(2), torch.zeros((), dtype=torch.long), encoder_hidden_states=ehs, return_dict=False)[0].squeeze(2)
    blocks = sorted([m for m in msa.torch_dfs(unet) if isinstance(m, attn_mod.TemporalBasicTransformerBlock)],
                    key=lambda m: -m.norm1.normalized_shape[0])
    assert len(blocks) == 16 and all(len(b.bank) == 1 for b in blocks)
    res = {f"bank{i:02d}": b.bank[0].numpy() for i, b in enumerate(blocks)}
    res["out"] = out.numpy()
    np.savez_compressed(os.path.join(GOLD, "refnet_tiny.npz"), **res)
    with open(os.path.join(GOLD, "refnet_spec.json"), "w") as f:
        json.dump([[k, list(s)] for k, s in spec], f)
    print("wrote refnet_tiny.npz:", [tuple(b.bank[0].shape) for b in blocks][:6], "...", tuple(out.shape), len(spec), "tensors")
    writer.clear()


if __name__ == "__main__":
    main()
