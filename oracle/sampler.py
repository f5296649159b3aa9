"""Context windows, DDIM (v-prediction, zero-SNR, trailing) and the denoise loop.  TEST INFRASTRUCTURE.

Restates src/pipelines/context.py:7-42, the diffusers-0.24 DDIMScheduler configuration of
config/prompts/animation.yaml:80-89 and the loop body of
Pose2VideoPipeline.__call__ (src/pipelines/pipeline_pose2vid_long.py:491-646).
"""
from typing import Callable, List

import numpy as np
import torch


def _bit_reversed_fraction(val: int) -> float:
    """context.py:7-12 ordered_halving."""
    return int(f"{val:064b}"[::-1], 2) / (1 << 64)


def uniform_windows(step: int, num_frames: int, context_size: int = 12, context_stride: int = 1,
                    context_overlap: int = 4, closed_loop: bool = True) -> List[List[int]]:
    """context.py:15-42."""
    if num_frames <= context_size:
        return [list(range(num_frames))]
    stride = min(context_stride, int(np.ceil(np.log2(num_frames / context_size))) + 1)
    frac = _bit_reversed_fraction(step)
    out = []
    for cstep in (1 << k for k in range(stride)):
        pad = int(round(num_frames * frac))
        start = int(frac * cstep) + pad
        stop = num_frames + pad + (0 if closed_loop else -context_overlap)
        for j in range(start, stop, context_size * cstep - context_overlap):
            out.append([e % num_frames for e in range(j, j + context_size * cstep, cstep)])
    return out


class DDIM:
    """DDIMScheduler(beta 0.00085..0.012 linear, zero-SNR rescale, v-prediction, trailing, eta=0)."""

    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012):
        betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        sqrt_ab = torch.cumprod(1.0 - betas, 0).sqrt()
        s0, sT = sqrt_ab[0].clone(), sqrt_ab[-1].clone()
        sqrt_ab = (sqrt_ab - sT) * (s0 / (s0 - sT))
        ab = sqrt_ab ** 2
        alphas = torch.cat([ab[0:1], ab[1:] / ab[:-1]])
        self.alphas_cumprod = torch.cumprod(alphas, 0)
        self.T = num_train_timesteps

    def timesteps(self, n: int) -> List[int]:
        return [int(v) for v in (np.round(np.arange(self.T, 0, -self.T / n)) - 1).astype(np.int64)]

    def coefficients(self, t: int, n: int):
        """x_prev = cx * x + cv * v  (v-prediction, eta = 0)."""
        a_t = float(self.alphas_cumprod[t])
        tp = t - self.T // n
        a_p = float(self.alphas_cumprod[tp]) if tp >= 0 else 1.0
        sa, sb = a_t ** 0.5, (1 - a_t) ** 0.5
        pa, pb = a_p ** 0.5, (1 - a_p) ** 0.5
        return pa * sa + pb * sb, -pa * sb + pb * sa

    def step(self, v, t: int, x, n: int):
        a_t = self.alphas_cumprod[t]
        tp = t - self.T // n
        a_p = self.alphas_cumprod[tp] if tp >= 0 else torch.tensor(1.0)
        x0 = a_t.sqrt() * x - (1 - a_t).sqrt() * v
        eps = a_t.sqrt() * v + (1 - a_t).sqrt() * x
        return a_p.sqrt() * x0 + (1 - a_p).sqrt() * eps


def denoise_step(unet_fn: Callable, latents, t: int, n_steps: int, ddim: DDIM, guidance_scale: float,
                 windows: List[List[int]], pose_fea, audio, full_mask, face_mask, lip_mask,
                 encoder_hidden_states, motion_scale):
    """One iteration of the loop at pipeline_pose2vid_long.py:494-635 (CFG on).

    latents (1,4,L,h,w); pose_fea (1,320,L,h,w); audio (2,L,32,768) = [zeros; audio];
    *_mask: list over levels of (2L, T_l) = cat([m]*2); encoder_hidden_states (2,1,768).
    unet_fn(sample, t, ehs, audio, pose, full, face, lip, motion_scale) -> (2,4,F,h,w).
    """
    L = latents.shape[2]
    noise = torch.zeros((2,) + tuple(latents.shape[1:]), dtype=latents.dtype)
    count = torch.zeros((1, 1, L, 1, 1), dtype=latents.dtype)
    for c in windows:
        x_in = latents[:, :, c].repeat(2, 1, 1, 1, 1)
        pose_in = pose_fea[:, :, c].repeat(2, 1, 1, 1, 1)
        aud_in = audio[:, c]
        gather = lambda ms: [m.view(2, L, -1)[:, c, :].reshape(-1, m.shape[-1]) for m in ms]  # noqa: E731
        pred = unet_fn(x_in, t, encoder_hidden_states, aud_in, pose_in, gather(full_mask),
                       gather(face_mask), gather(lip_mask), motion_scale)
        noise[:, :, c] = noise[:, :, c] + pred
        count[:, :, c] = count[:, :, c] + 1
    u, cnd = (noise / count).chunk(2)
    v = u + guidance_scale * (cnd - u)
    return ddim.step(v, t, latents, n_steps), v
