"""DDPMScheduler with the interface of monai-generative's `generative.networks.schedulers.DDPMScheduler`
(the external dependency the reference drives at train_ldm.py:74,145,160,163-167,351 and
train_ddpm.py:380-382), backed by fused sm_100a kernels:

  add_noise / get_velocity : one 128-bit vectorised kernel, per-sample coefficient gather (K14)
  step                     : ONE kernel per reverse step instead of ~12 elementwise launches (K15)

Index arithmetic (timesteps, alpha_bar[t], alpha_bar[t-1]) is done exactly as upstream: fp32 tables built
with torch on the host, `set_timesteps` in numpy int64, per-step coefficients evaluated in fp32 on the
host from those tables and passed to the kernel as scalars.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import ops
from ._lib import call


def _betas(schedule: str, num_train_timesteps: int, beta_start: float = 1e-4, beta_end: float = 2e-2,
           sig_range: float = 6.0, **unused) -> torch.Tensor:
    if schedule == "linear_beta":
        return torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
    if schedule == "scaled_linear_beta":
        return torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
    if schedule == "sigmoid_beta":
        return torch.sigmoid(torch.linspace(-sig_range, sig_range, num_train_timesteps)) * (beta_end - beta_start) \
            + beta_start
    raise ValueError(f"unknown noise schedule '{schedule}' (linear_beta, scaled_linear_beta, sigmoid_beta)")


class DDPMScheduler:
    def __init__(self, num_train_timesteps: int = 1000, schedule: str = "linear_beta",
                 variance_type: str = "fixed_small", clip_sample: bool = True, prediction_type: str = "epsilon",
                 **schedule_args) -> None:
        if variance_type not in ("fixed_small", "fixed_large", "learned", "learned_range"):
            raise ValueError("Argument `variance_type` must be a member of `DDPMVarianceType`")
        if prediction_type not in ("epsilon", "sample", "v_prediction"):
            raise ValueError("Argument `prediction_type` must be a member of `DDPMPredictionType`")
        if variance_type in ("learned", "learned_range"):
            raise NotImplementedError("learned variance types are not on the medimgen hot path")
        self.betas = _betas(schedule, num_train_timesteps, **schedule_args)
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.num_train_timesteps = num_train_timesteps
        self.one = torch.tensor(1.0)
        self.variance_type = variance_type
        self.clip_sample = clip_sample
        self.prediction_type = prediction_type
        self.num_inference_steps = None
        self.timesteps = torch.arange(num_train_timesteps - 1, -1, -1)
        self.noise_mode = "host"  # "host": z drawn on the CPU like upstream (seed-compatible); "device": on the GPU
        self._dev_tables: dict = {}

    # -- tables on the device (fp32, read by the gather in the add_noise kernel) -----------------------
    def _acp_on(self, device) -> torch.Tensor:
        t = self._dev_tables.get(device)
        if t is None:
            t = self.alphas_cumprod.to(device=device, dtype=torch.float32).contiguous()
            self._dev_tables[device] = t
        return t

    def _noise_op(self, a, b, timesteps, velocity: int):
        if not a.is_cuda:
            raise RuntimeError("DDPMScheduler: tensors must live on a CUDA device (no CPU path)")
        if a.dtype not in (torch.float32, torch.bfloat16):
            a = a.float()
        b = b.to(a.dtype)
        if a.stride() != b.stride() or not (a.is_contiguous() or ops._is_cl(a)):
            a, b = a.contiguous(), b.contiguous()
        B = a.shape[0]
        ts = timesteps.to(device=a.device, dtype=torch.int64).contiguous()
        if ts.numel() != B:
            raise RuntimeError(f"timesteps has {ts.numel()} entries for a batch of {B}")
        out = torch.empty_like(a)
        call("mig_ddpm_add_noise", ops._dt(a), ops._ptr(a), ops._ptr(b), ops._ptr(ts), ops._ptr(self._acp_on(a.device)),
             ops._ptr(out), B, a.numel() // B, self.num_train_timesteps, velocity, ops._stream())
        return out

    def add_noise(self, original_samples, noise, timesteps):
        """sqrt(acp[t]) * x0 + sqrt(1 - acp[t]) * noise, t per sample."""
        return self._noise_op(original_samples, noise, timesteps, 0)

    def get_velocity(self, sample, noise, timesteps):
        """sqrt(acp[t]) * noise - sqrt(1 - acp[t]) * sample."""
        return self._noise_op(sample, noise, timesteps, 1)

    # -- sampling ------------------------------------------------------------------------------------
    def set_timesteps(self, num_inference_steps: int, device=None) -> None:
        if num_inference_steps > self.num_train_timesteps:
            raise ValueError(f"`num_inference_steps`: {num_inference_steps} cannot be larger than "
                             f"`self.num_train_timesteps`: {self.num_train_timesteps}")
        self.num_inference_steps = num_inference_steps
        step_ratio = self.num_train_timesteps // self.num_inference_steps
        ts = (np.arange(0, num_inference_steps) * step_ratio).round()[::-1].astype(np.int64)
        self.timesteps = torch.from_numpy(ts.copy()).to(device)

    def step_coefficients(self, timestep: int) -> dict:
        """fp32 host evaluation of the per-step scalars, same operation order as upstream `step`."""
        t = int(timestep)
        acp_t = self.alphas_cumprod[t]
        acp_prev = self.alphas_cumprod[t - 1] if t > 0 else self.one
        beta_prod_t = 1 - acp_t
        beta_prod_prev = 1 - acp_prev
        c0 = (acp_prev ** 0.5 * self.betas[t]) / beta_prod_t
        ct = self.alphas[t] ** 0.5 * beta_prod_prev / beta_prod_t
        sigma = torch.tensor(0.0)
        if t > 0:
            var = (1 - acp_prev) / (1 - acp_t) * self.betas[t]
            if self.variance_type == "fixed_small":
                var = torch.clamp(var, min=1e-20)
            elif self.variance_type == "fixed_large":
                var = self.betas[t]
            sigma = var ** 0.5
        return dict(sqrt_acp=float(acp_t ** 0.5), sqrt_one_minus_acp=float(beta_prod_t ** 0.5), c0=float(c0),
                    ct=float(ct), sigma=float(sigma), t=t, t_prev=t - 1)

    def step(self, model_output, timestep, sample, generator=None, noise=None):
        """One reverse-diffusion step. Returns (pred_prev_sample, pred_original_sample).
        `noise` (optional) injects z for parity tests; otherwise z ~ N(0,1) is drawn per `noise_mode`."""
        if not sample.is_cuda:
            raise RuntimeError("DDPMScheduler.step: tensors must live on a CUDA device (no CPU path)")
        k = self.step_coefficients(int(timestep))
        x = sample if sample.dtype in (torch.float32, torch.bfloat16) else sample.float()
        eps = model_output.to(x.dtype)
        if eps.stride() != x.stride() or not (x.is_contiguous() or ops._is_cl(x)):
            x, eps = x.contiguous(), eps.contiguous()
        z = None
        if k["t"] > 0:
            if noise is not None:
                z = noise.to(device=x.device, dtype=x.dtype)
            elif self.noise_mode == "host":
                z = torch.randn(model_output.size(), dtype=model_output.dtype, layout=model_output.layout,
                                generator=generator).to(device=x.device, dtype=x.dtype)
            else:
                z = torch.randn(x.shape, dtype=x.dtype, device=x.device, generator=generator)
            if z.stride() != x.stride():
                z = z.contiguous() if x.is_contiguous() else ops.as_cl(z)
        prev = torch.empty_like(x)
        x0 = torch.empty_like(x)
        pred = {"epsilon": 0, "sample": 1, "v_prediction": 2}[self.prediction_type]
        call("mig_ddpm_step", ops._dt(x), ops._ptr(eps), ops._ptr(x), ops._ptr(z), ops._ptr(prev), ops._ptr(x0),
             x.numel(), k["sqrt_acp"], k["sqrt_one_minus_acp"], k["c0"], k["ct"], k["sigma"] if z is not None else 0.0,
             pred, int(self.clip_sample), ops._stream())
        return prev, x0
