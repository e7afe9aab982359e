"""DDIMScheduler (interface of generative/networks/schedulers/ddim.py) — step = one fused kernel launch."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from ... import _lib
from .scheduler import PRED_CODES, Scheduler, StrEnum, _f, _prep, _stream
from .._holders import on_input_device


class DDIMPredictionType(StrEnum):
    EPSILON = "epsilon"
    SAMPLE = "sample"
    V_PREDICTION = "v_prediction"


class DDIMScheduler(Scheduler):
    """ddim.py:55-301."""

    def __init__(self, num_train_timesteps: int = 1000, schedule: str = "linear_beta", clip_sample: bool = True,
                 set_alpha_to_one: bool = True, steps_offset: int = 0,
                 prediction_type: str = DDIMPredictionType.EPSILON, clip_sample_min: int = -1,
                 clip_sample_max: int = 1, **schedule_args) -> None:
        super().__init__(num_train_timesteps, schedule, **schedule_args)
        if prediction_type not in DDIMPredictionType.__members__.values():
            raise ValueError("Argument `prediction_type` must be a member of DDIMPredictionType")
        if clip_sample_min >= clip_sample_max:
            raise ValueError("clip_sample_min must be < clip_sample_max")
        self.prediction_type = prediction_type
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.first_alpha_cumprod = torch.tensor(0.0) if set_alpha_to_one else self.alphas_cumprod[-1]
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].astype(np.int64))
        self.clip_sample = clip_sample
        self.clip_sample_values = [clip_sample_min, clip_sample_max]
        self.steps_offset = steps_offset
        self.set_timesteps(self.num_train_timesteps)

    def set_timesteps(self, num_inference_steps: int, device: str | torch.device | None = None) -> None:
        self._check_steps(num_inference_steps)
        self.num_inference_steps = num_inference_steps
        step_ratio = self.num_train_timesteps // self.num_inference_steps
        timesteps = (np.arange(0, num_inference_steps) * step_ratio).round()[::-1].copy().astype(np.int64)
        self.timesteps = torch.from_numpy(timesteps).to(device)
        self.timesteps += self.steps_offset

    def _get_variance(self, timestep: int, prev_timestep: int) -> torch.Tensor:
        a_t = self.alphas_cumprod[timestep]
        a_prev = self.alphas_cumprod[prev_timestep] if prev_timestep >= 0 else self.final_alpha_cumprod
        return ((1 - a_prev) / (1 - a_t)) * (1 - a_t / a_prev)

    def _launch(self, model_output, sample, a_t, a_other, dir_sq, sigma, noise):
        lib = _lib.require_device()
        m, s, nz = _prep(model_output, sample, noise)
        c = _lib.DdimCoef()
        c.sqrt_alpha_prod_t, c.sqrt_beta_prod_t = _f(a_t ** 0.5), _f((1 - a_t) ** 0.5)
        c.sqrt_alpha_prod_prev, c.dir_coef, c.sigma = _f(a_other ** 0.5), _f(dir_sq ** 0.5), _f(sigma)
        c.clip_min, c.clip_max = float(self.clip_sample_values[0]), float(self.clip_sample_values[1])
        c.prediction_type, c.clip = PRED_CODES[str(self.prediction_type)], int(bool(self.clip_sample))
        prev, x0 = torch.empty_like(s), torch.empty_like(s)
        _lib.check(lib.b200_ddim_step(m.data_ptr(), s.data_ptr(), None if nz is None else nz.data_ptr(), C.byref(c),
                                      prev.data_ptr(), x0.data_ptr(), s.numel(), _stream()), "b200_ddim_step")
        if sample.dtype != torch.float32:
            prev, x0 = prev.to(sample.dtype), x0.to(sample.dtype)
        return prev, x0

    @on_input_device
    def step(self, model_output: torch.Tensor, timestep: int, sample: torch.Tensor, eta: float = 0.0,
             generator: torch.Generator | None = None) -> tuple[torch.Tensor, torch.Tensor]:
        """ddim.py:156-237 -> (pred_prev_sample, pred_original_sample)."""
        timestep = int(timestep)
        prev_timestep = timestep - self.num_train_timesteps // self.num_inference_steps
        a_t = self.alphas_cumprod[timestep]
        a_prev = self.alphas_cumprod[prev_timestep] if prev_timestep >= 0 else self.final_alpha_cumprod
        variance = self._get_variance(timestep, prev_timestep)
        std_dev_t = eta * variance ** 0.5
        noise = None
        sigma = 0.0
        if eta > 0:
            # drawn on the CPU with the caller's generator, exactly like the reference (ddim.py:229-235)
            noise = torch.randn(model_output.shape, dtype=model_output.dtype, generator=generator).to(model_output.device)
            sigma = variance ** 0.5 * eta
        return self._launch(model_output, sample, a_t, a_prev, 1 - a_prev - std_dev_t ** 2, sigma, noise)

    @on_input_device
    def reversed_step(self, model_output: torch.Tensor, timestep: int,
                      sample: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
        """ddim.py:239-301 -> (pred_next_sample, pred_original_sample)."""
        timestep = int(timestep)
        next_timestep = timestep + self.num_train_timesteps // self.num_inference_steps
        a_t = self.alphas_cumprod[timestep]
        a_next = (self.alphas_cumprod[next_timestep] if next_timestep < len(self.alphas_cumprod)
                  else self.first_alpha_cumprod)
        return self._launch(model_output, sample, a_t, a_next, 1 - a_next, 0.0, None)
