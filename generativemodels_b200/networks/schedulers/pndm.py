"""PNDMScheduler (interface of generative/networks/schedulers/pndm.py).

The Runge-Kutta accumulation, the Adams-Bashforth combine of up to four stored model outputs and formula (9) are a
single fused kernel (b200_pndm_step): it reads the history tensors and the sample once and writes x_{t-1} once.
The Python-side state the reference keeps across calls (``ets``, ``counter``, ``cur_sample``, ``cur_model_output``;
pndm.py:109-113) is kept with the same names and reset by ``set_timesteps`` only.
"""
from __future__ import annotations

import ctypes as C
from typing import Any

import numpy as np
import torch

from ... import _lib
from .scheduler import PRED_CODES, Scheduler, StrEnum, _f, _prep, _stream
from .._holders import on_input_device


class PNDMPredictionType(StrEnum):
    EPSILON = "epsilon"
    V_PREDICTION = "v_prediction"


class PNDMScheduler(Scheduler):
    """pndm.py:55-317."""

    def __init__(self, num_train_timesteps: int = 1000, schedule: str = "linear_beta", skip_prk_steps: bool = False,
                 set_alpha_to_one: bool = False, prediction_type: str = PNDMPredictionType.EPSILON,
                 steps_offset: int = 0, **schedule_args) -> None:
        super().__init__(num_train_timesteps, schedule, **schedule_args)
        if prediction_type not in PNDMPredictionType.__members__.values():
            raise ValueError("Argument `prediction_type` must be a member of PNDMPredictionType")
        self.prediction_type = prediction_type
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.init_noise_sigma = 1.0
        self.pndm_order = 4
        self.skip_prk_steps = skip_prk_steps
        self.steps_offset = steps_offset
        self.cur_model_output = 0
        self.counter = 0
        self.cur_sample = None
        self.ets: list = []
        self.set_timesteps(num_train_timesteps)

    def set_timesteps(self, num_inference_steps: int, device: str | torch.device | None = None) -> None:
        self._check_steps(num_inference_steps)
        self.num_inference_steps = num_inference_steps
        step_ratio = self.num_train_timesteps // self.num_inference_steps
        self._timesteps = (np.arange(0, num_inference_steps) * step_ratio).round().astype(np.int64)
        self._timesteps += self.steps_offset
        if self.skip_prk_steps:
            self.prk_timesteps = np.array([])
            self.plms_timesteps = self._timesteps[::-1]
        else:
            prk = np.array(self._timesteps[-self.pndm_order:]).repeat(2) + np.tile(
                np.array([0, self.num_train_timesteps // num_inference_steps // 2]), self.pndm_order)
            self.prk_timesteps = (prk[:-1].repeat(2)[1:-1])[::-1].copy()
            self.plms_timesteps = self._timesteps[:-3][::-1].copy()
        timesteps = np.concatenate([self.prk_timesteps, self.plms_timesteps]).astype(np.int64)
        self.timesteps = torch.from_numpy(timesteps).to(device)
        self.num_inference_steps = len(self.timesteps)
        self.ets = []
        self.counter = 0

    # -- fused kernel front-end --------------------------------------------------------------------
    def _combine(self, terms, sample=None, timestep=None, prev_timestep=None, want_eps=False):
        """eps = sum w_i * tensor_i; optionally x_prev = formula (9)(sample, eps) (pndm.py:293-315)."""
        lib = _lib.require_device()
        tens = _prep(*[t for _, t in terms])
        c = _lib.PndmCoef()
        c.n_hist = len(terms)
        for i, (w, _) in enumerate(terms):
            c.w[i] = float(w)
        c.prediction_type = PRED_CODES[str(self.prediction_type)]
        prev = None
        s = None
        if sample is not None:
            (s,) = _prep(sample)
            a_t = self.alphas_cumprod[timestep]
            a_prev = self.alphas_cumprod[prev_timestep] if prev_timestep >= 0 else self.final_alpha_cumprod
            b_t, b_prev = 1 - a_t, 1 - a_prev
            c.v_alpha, c.v_beta = _f(a_t ** 0.5), _f(b_t ** 0.5)
            c.sample_coeff = _f((a_prev / a_t) ** 0.5)
            denom = a_t * b_prev ** 0.5 + (a_t * b_t * a_prev) ** 0.5
            c.eps_coeff = _f((a_prev - a_t) / denom)
            prev = torch.empty_like(s)
        eps = torch.empty_like(tens[0]) if want_eps else None
        hist = (C.c_void_p * len(tens))(*[t.data_ptr() for t in tens])
        _lib.check(lib.b200_pndm_step(hist, None if s is None else s.data_ptr(), C.byref(c),
                                      None if prev is None else prev.data_ptr(),
                                      None if eps is None else eps.data_ptr(), tens[0].numel(), _stream()),
                   "b200_pndm_step")
        if prev is not None and sample.dtype != torch.float32:
            prev = prev.to(sample.dtype)
        return prev, eps

    @on_input_device
    def step(self, model_output: torch.Tensor, timestep: int, sample: torch.Tensor) -> tuple[torch.Tensor, Any]:
        """pndm.py:164-183."""
        if self.counter < len(self.prk_timesteps) and not self.skip_prk_steps:
            return self.step_prk(model_output=model_output, timestep=timestep, sample=sample), None
        return self.step_plms(model_output=model_output, timestep=timestep, sample=sample), None

    @staticmethod
    def _keep(model_output: torch.Tensor) -> torch.Tensor:
        """History entries must own their storage: a CUDA-graph-replayed network returns the same output buffer
        every step, which the next replay overwrites (the four PLMS weights would then all multiply the current eps)."""
        return model_output.clone()

    def step_prk(self, model_output: torch.Tensor, timestep: int, sample: torch.Tensor) -> torch.Tensor:
        """pndm.py:185-228 (Runge-Kutta warm-up, four network evaluations per step)."""
        if self.num_inference_steps is None:
            raise ValueError("Number of inference steps is 'None', you need to run 'set_timesteps' after creating "
                             "the scheduler")
        timestep = int(timestep)
        diff_to_prev = 0 if self.counter % 2 else self.num_train_timesteps // self.num_inference_steps // 2
        prev_timestep = timestep - diff_to_prev
        timestep = int(self.prk_timesteps[self.counter // 4 * 4])
        phase = self.counter % 4
        # like the reference, phase 0 ACCUMULATES (``self.cur_model_output += 1/6 * model_output``, pndm.py:208-209) into
        # whatever the last cycle left — 0 after a completed cycle, a stale partial sum after a loop aborted mid-cycle
        # (set_timesteps resets ``ets`` / ``counter`` only, pndm.py:160-161).  Kept bug-for-bug: results must be the
        # reference's on the same call sequence (tests/test_modules_cpu.py::test_pndm_restart_mid_runge_kutta_cpu).
        acc = self.cur_model_output
        has_acc = torch.is_tensor(acc)
        if phase == 0:
            self.ets.append(self._keep(model_output))
            self.cur_sample = sample
        if phase in (0, 1, 2):
            w = 1 / 6 if phase == 0 else 1 / 3
            terms_acc = ([(1.0, acc)] if has_acc else []) + [(w, model_output)]
            _, self.cur_model_output = self._combine(terms_acc, want_eps=True)
            terms = [(1.0, model_output)]
        else:
            terms = ([(1.0, acc)] if has_acc else []) + [(1 / 6, model_output)]
            self.cur_model_output = 0
        cur_sample = self.cur_sample if self.cur_sample is not None else sample
        prev_sample, _ = self._combine(terms, cur_sample, timestep, prev_timestep)
        self.counter += 1
        return prev_sample

    def step_plms(self, model_output: torch.Tensor, timestep: int, sample: torch.Tensor) -> torch.Tensor:
        """pndm.py:230-291 (linear multistep with the Adams-Bashforth weights)."""
        if self.num_inference_steps is None:
            raise ValueError("Number of inference steps is 'None', you need to run 'set_timesteps' after creating "
                             "the scheduler")
        if not self.skip_prk_steps and len(self.ets) < 3:
            raise ValueError(f"{self.__class__} can only be run AFTER scheduler has been run in 'prk' mode for at "
                             "least 12 iterations ")
        timestep = int(timestep)
        prev_timestep = timestep - self.num_train_timesteps // self.num_inference_steps
        if self.counter != 1:
            self.ets = self.ets[-3:]
            self.ets.append(self._keep(model_output))
        else:
            prev_timestep = timestep
            timestep = timestep + self.num_train_timesteps // self.num_inference_steps
        e = self.ets
        if len(e) == 1 and self.counter == 0:
            terms = [(1.0, model_output)]
            self.cur_sample = sample
        elif len(e) == 1 and self.counter == 1:
            terms = [(0.5, model_output), (0.5, e[-1])]
            sample = self.cur_sample
            self.cur_sample = None
        elif len(e) == 2:
            terms = [(3 / 2, e[-1]), (-1 / 2, e[-2])]
        elif len(e) == 3:
            terms = [(23 / 12, e[-1]), (-16 / 12, e[-2]), (5 / 12, e[-3])]
        else:
            terms = [(55 / 24, e[-1]), (-59 / 24, e[-2]), (37 / 24, e[-3]), (-9 / 24, e[-4])]
        prev_sample, _ = self._combine(terms, sample, timestep, prev_timestep)
        self.counter += 1
        return prev_sample
