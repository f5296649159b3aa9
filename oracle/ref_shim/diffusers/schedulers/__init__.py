"""DDIMScheduler restated from published diffusers 0.24.0 behaviour."""
from dataclasses import dataclass

import numpy as np
import torch


@dataclass
class DDIMSchedulerOutput:
    prev_sample: torch.Tensor
    pred_original_sample: torch.Tensor = None


def rescale_zero_terminal_snr(betas):
    alphas = 1.0 - betas
    alphas_cumprod = torch.cumprod(alphas, dim=0)
    s = alphas_cumprod.sqrt()
    s0, sT = s[0].clone(), s[-1].clone()
    s = s - sT
    s = s * (s0 / (s0 - sT))
    ab = s ** 2
    a = ab[1:] / ab[:-1]
    a = torch.cat([ab[0:1], a])
    return 1 - a


class DDIMScheduler:
    order = 1

    def __init__(self, num_train_timesteps=1000, beta_start=0.0001, beta_end=0.02, beta_schedule="linear",
                 clip_sample=True, set_alpha_to_one=True, steps_offset=0, prediction_type="epsilon",
                 timestep_spacing="leading", rescale_betas_zero_snr=False, **unused):
        assert beta_schedule == "linear"
        self.num_train_timesteps = num_train_timesteps
        self.betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        if rescale_betas_zero_snr:
            self.betas = rescale_zero_terminal_snr(self.betas)
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.init_noise_sigma = 1.0
        self.clip_sample = clip_sample
        self.prediction_type = prediction_type
        self.timestep_spacing = timestep_spacing
        self.steps_offset = steps_offset

    def scale_model_input(self, sample, timestep=None):
        return sample

    def set_timesteps(self, n, device=None):
        self.num_inference_steps = n
        if self.timestep_spacing == "trailing":
            ts = np.round(np.arange(self.num_train_timesteps, 0, -self.num_train_timesteps / n)) - 1
        elif self.timestep_spacing == "leading":
            ts = (np.arange(0, n) * (self.num_train_timesteps // n)).round()[::-1].copy() + self.steps_offset
        else:
            raise ValueError(self.timestep_spacing)
        self.timesteps = torch.from_numpy(ts.astype(np.int64)).to(device)

    def step(self, model_output, timestep, sample, eta=0.0, generator=None, **kw):
        assert eta == 0.0 and not self.clip_sample
        t = int(timestep)
        prev_t = t - self.num_train_timesteps // self.num_inference_steps
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.final_alpha_cumprod
        b_t = 1 - a_t
        if self.prediction_type == "v_prediction":
            x0 = (a_t ** 0.5) * sample - (b_t ** 0.5) * model_output
            eps = (a_t ** 0.5) * model_output + (b_t ** 0.5) * sample
        elif self.prediction_type == "epsilon":
            x0 = (sample - b_t ** 0.5 * model_output) / a_t ** 0.5
            eps = model_output
        else:
            raise ValueError(self.prediction_type)
        direction = (1 - a_prev) ** 0.5 * eps
        prev = a_prev ** 0.5 * x0 + direction
        return DDIMSchedulerOutput(prev_sample=prev, pred_original_sample=x0)


DPMSolverMultistepScheduler = EulerAncestralDiscreteScheduler = EulerDiscreteScheduler = None
LMSDiscreteScheduler = PNDMScheduler = None
