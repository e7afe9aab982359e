"""DDPMScheduler (interface of generative/networks/schedulers/ddpm.py) — step = one fused kernel launch."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from ... import _lib
from .scheduler import PRED_CODES, Scheduler, StrEnum, _f, _prep, _stream
from .._holders import on_input_device


class DDPMPredictionType(StrEnum):
    EPSILON = "epsilon"
    SAMPLE = "sample"
    V_PREDICTION = "v_prediction"


class DDPMVarianceType(StrEnum):
    FIXED_SMALL = "fixed_small"
    FIXED_LARGE = "fixed_large"
    LEARNED = "learned"
    LEARNED_RANGE = "learned_range"


class DDPMScheduler(Scheduler):
    """ddpm.py:66-252."""

    def __init__(self, num_train_timesteps: int = 1000, schedule: str = "linear_beta",
                 variance_type: str = DDPMVarianceType.FIXED_SMALL, clip_sample: bool = True,
                 prediction_type: str = DDPMPredictionType.EPSILON, clip_sample_min: int = -1,
                 clip_sample_max: int = 1, **schedule_args) -> None:
        super().__init__(num_train_timesteps, schedule, **schedule_args)
        if variance_type not in DDPMVarianceType.__members__.values():
            raise ValueError("Argument `variance_type` must be a member of `DDPMVarianceType`")
        if prediction_type not in DDPMPredictionType.__members__.values():
            raise ValueError("Argument `prediction_type` must be a member of `DDPMPredictionType`")
        if clip_sample_min >= clip_sample_max:
            raise ValueError("clip_sample_min must be < clip_sample_max")
        self.clip_sample = clip_sample
        self.clip_sample_values = [clip_sample_min, clip_sample_max]
        self.variance_type = variance_type
        self.prediction_type = prediction_type

    def set_timesteps(self, num_inference_steps: int, device: str | torch.device | None = None) -> None:
        self._check_steps(num_inference_steps)
        self.num_inference_steps = num_inference_steps
        step_ratio = self.num_train_timesteps // self.num_inference_steps
        timesteps = (np.arange(0, num_inference_steps) * step_ratio).round()[::-1].astype(np.int64)
        self.timesteps = torch.from_numpy(timesteps.copy()).to(device)

    def _get_mean(self, timestep: int, x_0: torch.Tensor, x_t: torch.Tensor) -> torch.Tensor:
        """ddpm.py:133-156 (likelihood path; plain tensor expression on the caller's device)."""
        alpha_t = self.alphas[timestep]
        a_t = self.alphas_cumprod[timestep]
        a_prev = self.alphas_cumprod[timestep - 1] if timestep > 0 else self.one
        c0 = a_prev.sqrt() * self.betas[timestep] / (1 - a_t)
        ct = alpha_t.sqrt() * (1 - a_prev) / (1 - a_t)
        return c0 * x_0 + ct * x_t

    def _get_variance(self, timestep: int, predicted_variance: torch.Tensor | None = None) -> torch.Tensor:
        """ddpm.py:158-189."""
        a_t = self.alphas_cumprod[timestep]
        a_prev = self.alphas_cumprod[timestep - 1] if timestep > 0 else self.one
        variance = (1 - a_prev) / (1 - a_t) * self.betas[timestep]
        if self.variance_type == DDPMVarianceType.FIXED_SMALL:
            variance = torch.clamp(variance, min=1e-20)
        elif self.variance_type == DDPMVarianceType.FIXED_LARGE:
            variance = self.betas[timestep]
        elif self.variance_type == DDPMVarianceType.LEARNED:
            return predicted_variance
        elif self.variance_type == DDPMVarianceType.LEARNED_RANGE:
            frac = (predicted_variance + 1) / 2
            variance = frac * self.betas[timestep] + (1 - frac) * variance
        return variance

    @on_input_device
    def step(self, model_output: torch.Tensor, timestep: int, sample: torch.Tensor,
             generator: torch.Generator | None = None) -> tuple[torch.Tensor, torch.Tensor]:
        """ddpm.py:191-252 -> (pred_prev_sample, pred_original_sample)."""
        lib = _lib.require_device()
        timestep = int(timestep)
        predicted_variance = None
        if model_output.shape[1] == sample.shape[1] * 2 and self.variance_type in ["learned", "learned_range"]:
            model_output, predicted_variance = torch.split(model_output, sample.shape[1], dim=1)
        a_t = self.alphas_cumprod[timestep]
        a_prev = self.alphas_cumprod[timestep - 1] if timestep > 0 else self.one
        b_t, b_prev = 1 - a_t, 1 - a_prev
        c = _lib.DdpmCoef()
        c.sqrt_alpha_prod_t, c.sqrt_beta_prod_t = _f(a_t ** 0.5), _f(b_t ** 0.5)
        c.coef_x0 = _f((a_prev ** 0.5 * self.betas[timestep]) / b_t)
        c.coef_xt = _f(self.alphas[timestep] ** 0.5 * b_prev / b_t)
        c.clip_min, c.clip_max = float(self.clip_sample_values[0]), float(self.clip_sample_values[1])
        c.prediction_type, c.clip = PRED_CODES[str(self.prediction_type)], int(bool(self.clip_sample))
        noise = None
        if timestep > 0:
            # CPU draw + copy, like the reference (ddpm.py:245-247), so seeded runs see identical noise
            noise = torch.randn(model_output.size(), dtype=model_output.dtype, layout=model_output.layout,
                                generator=generator).to(model_output.device)
            if predicted_variance is not None and self.variance_type == DDPMVarianceType.LEARNED:
                c.var_mode = 1
            elif predicted_variance is not None and self.variance_type == DDPMVarianceType.LEARNED_RANGE:
                c.var_mode = 2
                c.min_log = _f((1 - a_prev) / (1 - a_t) * self.betas[timestep])
                c.max_log = _f(self.betas[timestep])
            else:
                c.var_mode = 0
                c.sigma = _f(self._get_variance(timestep) ** 0.5)
        m, s, nz, pv = _prep(model_output, sample, noise, predicted_variance if noise is not None else None)
        prev, x0 = torch.empty_like(s), torch.empty_like(s)
        _lib.check(lib.b200_ddpm_step(m.data_ptr(), s.data_ptr(), None if nz is None else nz.data_ptr(),
                                      None if pv is None else pv.data_ptr(), C.byref(c), prev.data_ptr(), x0.data_ptr(),
                                      s.numel(), _stream()), "b200_ddpm_step")
        if sample.dtype != torch.float32:
            prev, x0 = prev.to(sample.dtype), x0.to(sample.dtype)
        return prev, x0
