"""Seeded synthetic inputs for the hot path (SURVEY.md section 8d "Synthetic inputs").  TEST INFRASTRUCTURE.

Everything is generated on the CPU with explicit generators so the reference, the oracle and the
CUDA path consume bit-identical inputs on any machine with this image.
"""
from typing import Dict, List

import numpy as np
import torch

from .mask_pyramid import full_mask_from_lips, preprocess_mov_mask
from .unet3d import UNetSpec, spatial_block_prefixes, spatial_block_width


def _box_blur_u8(img: np.ndarray, k: int) -> np.ndarray:
    """Cheap separable blur + min-max normalise to 0..255 (stands in for cv2.GaussianBlur +
    cv2.normalize of src/utils/util.py:19-39; only the *shape* of the data matters here)."""
    x = img.astype(np.float32)
    ker = np.ones(k, dtype=np.float32) / k
    x = np.apply_along_axis(lambda r: np.convolve(r, ker, mode="same"), -1, x)
    x = np.apply_along_axis(lambda r: np.convolve(r, ker, mode="same"), -2, x)
    lo, hi = x.min(), x.max()
    if hi > lo:
        x = (x - lo) / (hi - lo) * 255.0
    return np.clip(np.rint(x), 0, 255).astype(np.uint8)


def synthetic_masks_u8(num_frames: int, seed: int = 7, size: int = 64):
    """Per frame: 1-3 random rectangles, blurred, normalised -> (L, size, size) uint8 (face, lips)."""
    rng = np.random.default_rng(seed)
    out = []
    for kind, k in (("face", 9), ("lips", 5)):
        frames = np.zeros((num_frames, size, size), dtype=np.uint8)
        for f in range(num_frames):
            for _ in range(int(rng.integers(1, 4))):
                y0, x0 = rng.integers(0, size - 8, 2)
                hh, ww = rng.integers(4, size // 2, 2)
                frames[f, y0:y0 + hh, x0:x0 + ww] = 255
            frames[f] = _box_blur_u8(frames[f], k)
        out.append(frames)
    return out[0], out[1]


def level_tokens(latent: int) -> List[int]:
    return [(latent >> k) ** 2 for k in range(4)]


def make_inputs(spec: UNetSpec, video_length: int, latent: int, seed: int = 42) -> Dict:
    """Whole-video inputs in the layout Pose2VideoPipeline holds them right before the loop
    (pipeline_pose2vid_long.py:411-486), CFG on."""
    g = torch.Generator().manual_seed(seed)
    L = video_length
    c0 = spec.block_out_channels[0]
    latents = torch.randn(1, spec.in_channels, L, latent, latent, generator=g)
    clip = torch.randn(1, 1, spec.cross_attention_dim, generator=g)
    ehs = torch.cat([torch.zeros_like(clip), clip], dim=0)                       # :389-394
    aud = torch.randn(1, L, 32, spec.audio_attention_dim, generator=g)
    aud = torch.nn.functional.layer_norm(aud, (spec.audio_attention_dim,))
    audio = torch.cat([torch.zeros_like(aud), aud], dim=0)                       # :484-485
    pose = 0.1 * torch.randn(1, c0, L, latent, latent, generator=g)
    face_u8, lips_u8 = synthetic_masks_u8(L)
    face, lips = preprocess_mov_mask(face_u8, lips_u8, image_size=latent * 8)
    full = full_mask_from_lips(lips)
    dup = lambda ms: [torch.from_numpy(np.concatenate([m, m], 0)) for m in ms]   # noqa: E731  (:451-465)
    return dict(latents=latents, encoder_hidden_states=ehs, audio=audio, pose_fea=pose,
                full_mask=dup(full), face_mask=dup(face), lip_mask=dup(lips),
                face_u8=face_u8, lips_u8=lips_u8, motion_scale=[1.0, 1.0, 2.0])


def make_banks(spec: UNetSpec, latent: int, seed: int = 1234, batch: int = 2) -> Dict[str, torch.Tensor]:
    """16 reference-feature banks (Bb, T_l, C_l), rounded through fp16 as
    ReferenceAttentionControl.update does by default (mutual_self_attention.py:304,340)."""
    g = torch.Generator().manual_seed(seed)
    banks = {}
    for pre in spatial_block_prefixes(spec):
        c = spatial_block_width(spec, pre)
        parts = pre.split(".")
        if parts[0] == "down_blocks":
            lvl = int(parts[1])
        elif parts[0] == "mid_block":
            lvl = 3
        else:
            lvl = 3 - int(parts[1])
        t = (latent >> lvl) ** 2
        banks[pre] = torch.randn(batch, t, c, generator=g).to(torch.float16).to(torch.float32)
    return banks


def window_inputs(inp: Dict, frames: List[int]):
    """Gather one context window exactly like pipeline_pose2vid_long.py:556-586 (CFG on)."""
    L = inp["latents"].shape[2]
    g = lambda ms: [m.view(2, L, -1)[:, frames, :].reshape(-1, m.shape[-1]) for m in ms]  # noqa: E731
    return dict(sample=inp["latents"][:, :, frames].repeat(2, 1, 1, 1, 1),
                encoder_hidden_states=inp["encoder_hidden_states"],
                audio_embedding=inp["audio"][:, frames],
                pose_cond_fea=inp["pose_fea"][:, :, frames].repeat(2, 1, 1, 1, 1),
                full_mask=g(inp["full_mask"]), face_mask=g(inp["face_mask"]), body_mask=g(inp["lip_mask"]),
                motion_scale=inp["motion_scale"])
