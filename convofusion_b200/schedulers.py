"""Host mirrors of the diffusers==0.14.0 schedulers the reference instantiates through its
`target:` registry (configs/modules/scheduler.yaml; call sites convofusion.py:104-106,419-429,544-545
and unbounded_synthesis.py:49,56-58,75,181).

diffusers is a third-party dependency that is absent from the reference tree and from this image; the
classes below restate its published DDIM / DDPM algorithm with the same float32 table arithmetic and
the same public members the reference touches: `init_noise_sigma`, `set_timesteps`, `timesteps`,
`step(...).prev_sample / .pred_original_sample` (with `eta` in DDIM's signature only, which is how the
reference detects DDIM), `add_noise`, `betas`, `config.num_train_timesteps`.

Scalar work (beta tables, per-step coefficients) stays on the host -- it is O(steps); tensor work goes
through the fused guidance+step kernel (cfb_guidance_sched_step) or, inside Denoiser.sample(), the
graph-captured loop.  `step_table()` exports the per-step coefficient rows that kernel consumes.
"""
from __future__ import annotations

import ctypes as C
from types import SimpleNamespace
from typing import Optional

import hashlib

import numpy as np
import torch

from . import _lib


class SchedulerOutput:
    def __init__(self, prev_sample, pred_original_sample=None):
        self.prev_sample = prev_sample
        self.pred_original_sample = pred_original_sample


class _Base:
    kind = None

    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 0.0001, beta_end: float = 0.02,
                 beta_schedule: str = "linear", trained_betas=None, clip_sample: bool = True,
                 prediction_type: str = "epsilon", **extra):
        if trained_betas is not None:
            self.betas = torch.tensor(trained_betas, dtype=torch.float32)
        elif beta_schedule == "linear":
            self.betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        elif beta_schedule == "scaled_linear":
            self.betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps,
                                        dtype=torch.float32) ** 2
        else:
            raise NotImplementedError(f"{beta_schedule} does is not implemented for {self.__class__}")
        if prediction_type != "epsilon":
            raise NotImplementedError("only prediction_type='epsilon' (PREDICT_EPSILON: True) is implemented")
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.one = torch.tensor(1.0)
        self.init_noise_sigma = 1.0
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy())
        self.config = SimpleNamespace(num_train_timesteps=num_train_timesteps, beta_start=beta_start,
                                      beta_end=beta_end, beta_schedule=beta_schedule, clip_sample=clip_sample,
                                      prediction_type=prediction_type, **extra)

    def scale_model_input(self, sample, timestep=None):
        return sample

    def add_noise(self, original_samples, noise, timesteps):
        acp = self.alphas_cumprod.to(device=original_samples.device, dtype=original_samples.dtype)
        t = torch.as_tensor(timesteps).reshape(-1).long().to(original_samples.device)
        sa, sb = acp[t] ** 0.5, (1 - acp[t]) ** 0.5
        while sa.dim() < original_samples.dim():
            sa, sb = sa.unsqueeze(-1), sb.unsqueeze(-1)
        return sa * original_samples + sb * noise

    # ---- coefficient rows: {sqrt(1-abar_t), sqrt(abar_t), k0, k1, k2, ia, ib, 0}, float32 arithmetic
    def _row(self, t: int, eta: float):
        raise NotImplementedError

    def step_table(self, num_steps: int, eta: float = 0.0, noise_scheduler=None) -> dict:
        """Per-step coefficient rows for the fused loop.  O(steps) scalar work on the host (a few ms of torch scalar
        ops), so the table is memoised per (steps, eta, noise schedule): repeated sampling calls pay it once."""
        self.set_timesteps(num_steps)
        ns = noise_scheduler if noise_scheduler is not None else self
        # the noise schedule itself is part of the key: `trained_betas` is not in `config`, two schedulers with the same
        # config can carry different alphas_cumprod
        acp = ns.alphas_cumprod.detach().to("cpu", torch.float64).contiguous().numpy()
        key = (int(num_steps), float(eta), ns.kind, repr(sorted(vars(ns.config).items())),
               hashlib.sha1(acp.tobytes()).hexdigest())
        cache = self.__dict__.setdefault("_step_tables", {})
        if key in cache:
            return cache[key]
        ts = self.timesteps.cpu().numpy().astype(np.int64)
        coef = np.zeros((len(ts), 8), dtype=np.float32)
        for i, t in enumerate(ts):
            coef[i, :5] = self._row(int(t), eta)
            a = ns.alphas_cumprod[int(t)]
            coef[i, 5], coef[i, 6] = float(a ** 0.5), float((1 - a) ** 0.5)
        cache[key] = {"kind": self.kind, "timesteps": ts, "coef": coef, "clip_sample": bool(self.config.clip_sample),
                      "needs_noise": bool(np.any(coef[:, 4] != 0))}
        return cache[key]

    def _device_step(self, model_output, sample, row, noise):
        if model_output.device.type != "cuda":
            raise _lib.CfbError("scheduler.step needs CUDA tensors: convofusion_b200 has no CPU path")
        eps = model_output.detach().to(torch.float32).contiguous()
        x = sample.detach().to(torch.float32).contiguous().clone()
        if row[4] != 0 and noise is None:
            noise = torch.randn(eps.shape, device=eps.device, dtype=eps.dtype)   # global RNG like diffusers
        coef = torch.tensor(np.concatenate([row, np.zeros(3, np.float32)]), device=eps.device)
        n_clips = x.shape[0]
        with torch.cuda.device(eps.device):
            _lib.check(_lib.lib().cfb_guidance_sched_step(
                eps.data_ptr(), x.data_ptr(), _lib.ptr(noise.contiguous() if noise is not None else None),
                coef.data_ptr(), 1, n_clips, x.numel() // n_clips, self.kind, int(self.config.clip_sample), 1.0,
                _lib.stream_ptr()))
        return SchedulerOutput(x)


class DDIMScheduler(_Base):
    kind = _lib.SCHED_DDIM

    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 0.0001, beta_end: float = 0.02,
                 beta_schedule: str = "linear", trained_betas=None, clip_sample: bool = True,
                 set_alpha_to_one: bool = True, steps_offset: int = 0, prediction_type: str = "epsilon", **kw):
        super().__init__(num_train_timesteps, beta_start, beta_end, beta_schedule, trained_betas, clip_sample,
                         prediction_type, set_alpha_to_one=set_alpha_to_one, steps_offset=steps_offset, **kw)
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]

    def set_timesteps(self, num_inference_steps: int, device=None):
        if num_inference_steps > self.config.num_train_timesteps:
            raise ValueError(f"`num_inference_steps`: {num_inference_steps} cannot be larger than "
                             f"`self.config.train_timesteps`: {self.config.num_train_timesteps}")
        self.num_inference_steps = num_inference_steps
        ratio = self.config.num_train_timesteps // num_inference_steps
        ts = (np.arange(0, num_inference_steps) * ratio).round()[::-1].copy().astype(np.int64)
        self.timesteps = torch.from_numpy(ts).to(device) + self.config.steps_offset

    def _row(self, t: int, eta: float):
        if self.num_inference_steps is None:
            raise ValueError("Number of inference steps is 'None', you need to run 'set_timesteps' after creating the scheduler")
        prev = t - self.config.num_train_timesteps // self.num_inference_steps
        a_t = self.alphas_cumprod[t]
        a_p = self.alphas_cumprod[prev] if prev >= 0 else self.final_alpha_cumprod
        b_t = 1 - a_t
        variance = ((1 - a_p) / (1 - a_t)) * (1 - a_t / a_p)
        std = eta * variance ** 0.5
        k1 = (1 - a_p - std ** 2) ** 0.5
        k2 = std if eta > 0 else torch.tensor(0.0)
        return np.array([float(b_t ** 0.5), float(a_t ** 0.5), float(a_p ** 0.5), float(k1), float(k2)], np.float32)

    def step(self, model_output, timestep, sample, eta: float = 0.0, use_clipped_model_output: bool = False,
             generator=None, variance_noise=None, return_dict: bool = True):
        if use_clipped_model_output:
            raise NotImplementedError("use_clipped_model_output is not implemented")
        return self._device_step(model_output, sample, self._row(int(timestep), eta), variance_noise)


class DDPMScheduler(_Base):
    kind = _lib.SCHED_DDPM

    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 0.0001, beta_end: float = 0.02,
                 beta_schedule: str = "linear", trained_betas=None, variance_type: str = "fixed_small",
                 clip_sample: bool = True, prediction_type: str = "epsilon", **kw):
        if variance_type != "fixed_small":
            raise NotImplementedError("only variance_type='fixed_small' (configs/modules/scheduler.yaml) is implemented")
        super().__init__(num_train_timesteps, beta_start, beta_end, beta_schedule, trained_betas, clip_sample,
                         prediction_type, variance_type=variance_type, **kw)

    def set_timesteps(self, num_inference_steps: int, device=None):
        n = min(self.config.num_train_timesteps, num_inference_steps)
        self.num_inference_steps = n
        ts = np.arange(0, self.config.num_train_timesteps, self.config.num_train_timesteps // n)[::-1].copy()
        self.timesteps = torch.from_numpy(ts).to(device)

    def _row(self, t: int, eta: float = 0.0):
        n = self.num_inference_steps if self.num_inference_steps else self.config.num_train_timesteps
        prev = t - self.config.num_train_timesteps // n
        a_t = self.alphas_cumprod[t]
        a_p = self.alphas_cumprod[prev] if prev >= 0 else self.one
        b_t, b_p = 1 - a_t, 1 - a_p
        cur_a = a_t / a_p
        cur_b = 1 - cur_a
        c0 = (a_p ** 0.5 * cur_b) / b_t
        c1 = cur_a ** 0.5 * b_p / b_t
        sigma = torch.clamp(b_p / b_t * cur_b, min=1e-20) ** 0.5 if t > 0 else torch.tensor(0.0)
        return np.array([float(b_t ** 0.5), float(a_t ** 0.5), float(c0), float(c1), float(sigma)], np.float32)

    def step(self, model_output, timestep, sample, generator=None, variance_noise=None, return_dict: bool = True):
        return self._device_step(model_output, sample, self._row(int(timestep)), variance_noise)
