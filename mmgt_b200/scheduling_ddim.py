"""Host-side DDIM tables for the denoise loop (diffusers==0.24.0 DDIMScheduler as configured by
config/prompts/animation.yaml:80-89: linear betas 0.00085..0.012, zero-terminal-SNR rescale, v-prediction,
trailing timestep spacing, eta = 0).  Only scalars are computed here; the tensor update runs in
``mmgt_cfg_ddim_step``.  A diffusers scheduler object can be adopted with ``DDIMSchedule.from_scheduler``.
"""
from dataclasses import dataclass
from typing import List

import numpy as np
import torch


def _rescale_zero_terminal_snr(betas: torch.Tensor) -> torch.Tensor:
    sqrt_ab = torch.cumprod(1.0 - betas, 0).sqrt()
    s0, sT = sqrt_ab[0].clone(), sqrt_ab[-1].clone()
    sqrt_ab = (sqrt_ab - sT) * (s0 / (s0 - sT))
    ab = sqrt_ab ** 2
    alphas = torch.cat([ab[0:1], ab[1:] / ab[:-1]])
    return 1.0 - alphas


@dataclass
class DDIMSchedule:
    alphas_cumprod: torch.Tensor               # (T,) float32 on CPU
    num_train_timesteps: int = 1000
    prediction_type: str = "v_prediction"
    timestep_spacing: str = "trailing"
    steps_offset: int = 1
    final_alpha_cumprod: float = 1.0
    init_noise_sigma: float = 1.0

    @classmethod
    def from_config(cls, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="linear",
                    rescale_betas_zero_snr=True, prediction_type="v_prediction", timestep_spacing="trailing",
                    steps_offset=1, set_alpha_to_one=True, clip_sample=False, **unused):
        if beta_schedule != "linear" or clip_sample:
            raise NotImplementedError("only the animation.yaml scheduler configuration is implemented")
        betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        if rescale_betas_zero_snr:
            betas = _rescale_zero_terminal_snr(betas)
        ac = torch.cumprod(1.0 - betas, 0)
        return cls(ac, num_train_timesteps, prediction_type, timestep_spacing, steps_offset,
                   1.0 if set_alpha_to_one else float(ac[0]))

    @classmethod
    def from_scheduler(cls, scheduler):
        """Adopt a diffusers DDIMScheduler instance (what scripts/pose2vid.py:170 constructs)."""
        cfg = getattr(scheduler, "config", scheduler)
        get = lambda k, d: getattr(cfg, k, getattr(scheduler, k, d))  # noqa: E731
        return cls(scheduler.alphas_cumprod.detach().float().cpu(), int(get("num_train_timesteps", 1000)),
                   get("prediction_type", "v_prediction"), get("timestep_spacing", "trailing"),
                   int(get("steps_offset", 1)), float(getattr(scheduler, "final_alpha_cumprod", 1.0)))

    def timesteps(self, n: int) -> List[int]:
        T = self.num_train_timesteps
        if self.timestep_spacing == "trailing":
            ts = np.round(np.arange(T, 0, -T / n)) - 1
        elif self.timestep_spacing == "leading":
            ts = (np.arange(0, n) * (T // n)).round()[::-1].copy() + self.steps_offset
        else:
            raise NotImplementedError(self.timestep_spacing)
        return [int(v) for v in ts.astype(np.int64)]

    def step_coefficients(self, t: int, n: int):
        """x_prev = cx * x + cv * model_output for eta = 0."""
        a_t = float(self.alphas_cumprod[t])
        tp = t - self.num_train_timesteps // n
        a_p = float(self.alphas_cumprod[tp]) if tp >= 0 else self.final_alpha_cumprod
        sa, sb, pa, pb = a_t ** 0.5, (1 - a_t) ** 0.5, a_p ** 0.5, (1 - a_p) ** 0.5
        if self.prediction_type == "v_prediction":
            # x0 = sa*x - sb*v ; eps = sa*v + sb*x
            return pa * sa + pb * sb, pb * sa - pa * sb
        if self.prediction_type == "epsilon":
            # x0 = (x - sb*e)/sa ; eps = e
            return pa / sa, pb - pa * sb / sa
        raise NotImplementedError(self.prediction_type)
