"""CPU restatement of monai-generative's DDPMScheduler + Diffusion/LatentDiffusion inferers.

TEST INFRASTRUCTURE (see oracle/__init__.py). PARITY UNPINNED: the arithmetic lives in the
third-party package `monai-generative` (import name `generative`), listed unpinned at
/root/reference/pyproject.toml:23 and absent from /root/reference and from this image. This file
restates its published algorithm (Ho et al. DDPM posterior with MONAI-Generative's conventions)
and is anchored on the reference's own call sites:
  train_ldm.py:74 (constructor kwargs), :145 (num_train_timesteps), :160 (add_noise),
  :163-167 (prediction_type / get_velocity), :351 (set_timesteps), :362 (inferer.sample);
  train_ddpm.py:189-192 (inferer __call__), :240-244 (sample), :380-382 (kwargs).
Plain torch CPU fp32 + numpy integer arithmetic; no device code.
"""
from __future__ import annotations

import numpy as np
import torch


def make_betas(schedule: str, num_train_timesteps: int, beta_start: float = 1e-4, beta_end: float = 2e-2,
               **kw) -> torch.Tensor:
    """NoiseSchedules registry of generative.networks.schedulers.scheduler (fp32 tables)."""
    if schedule == "linear_beta":
        return torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
    if schedule == "scaled_linear_beta":  # the LDM default, configuration.py:1012-1013
        return torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
    if schedule == "sigmoid_beta":
        sig_range = kw.get("sig_range", 6.0)
        b = torch.linspace(-sig_range, sig_range, num_train_timesteps)
        return torch.sigmoid(b) * (beta_end - beta_start) + beta_start
    raise ValueError(f"unknown schedule {schedule}")


class OracleDDPMScheduler:
    def __init__(self, num_train_timesteps: int = 1000, schedule: str = "linear_beta",
                 variance_type: str = "fixed_small", clip_sample: bool = True,
                 prediction_type: str = "epsilon", **schedule_args):
        self.betas = make_betas(schedule, num_train_timesteps, **schedule_args)
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.num_train_timesteps = num_train_timesteps
        self.one = torch.tensor(1.0)
        self.variance_type = variance_type
        self.clip_sample = clip_sample
        self.prediction_type = prediction_type
        self.num_inference_steps = None
        self.timesteps = torch.arange(num_train_timesteps - 1, -1, -1)

    # -- training side -------------------------------------------------------------------------
    def add_noise(self, original_samples, noise, timesteps):
        acp = self.alphas_cumprod.to(dtype=original_samples.dtype)
        t = timesteps.long()
        shape = (-1,) + (1,) * (original_samples.ndim - 1)
        a = (acp[t] ** 0.5).reshape(shape)
        b = ((1 - acp[t]) ** 0.5).reshape(shape)
        return a * original_samples + b * noise

    def get_velocity(self, sample, noise, timesteps):
        acp = self.alphas_cumprod.to(dtype=sample.dtype)
        t = timesteps.long()
        shape = (-1,) + (1,) * (sample.ndim - 1)
        a = (acp[t] ** 0.5).reshape(shape)
        b = ((1 - acp[t]) ** 0.5).reshape(shape)
        return a * noise - b * sample

    # -- sampling side -------------------------------------------------------------------------
    def set_timesteps(self, num_inference_steps: int, device=None):
        if num_inference_steps > self.num_train_timesteps:
            raise ValueError("num_inference_steps cannot exceed num_train_timesteps")
        self.num_inference_steps = num_inference_steps
        step_ratio = self.num_train_timesteps // num_inference_steps
        ts = (np.arange(0, num_inference_steps) * step_ratio).round()[::-1].astype(np.int64)
        self.timesteps = torch.from_numpy(ts.copy())

    def _variance(self, t: int):
        acp_t = self.alphas_cumprod[t]
        acp_prev = self.alphas_cumprod[t - 1] if t > 0 else self.one
        var = (1 - acp_prev) / (1 - acp_t) * self.betas[t]
        if self.variance_type == "fixed_small":
            var = torch.clamp(var, min=1e-20)
        elif self.variance_type == "fixed_large":
            var = self.betas[t]
        else:
            raise ValueError("learned variance types are not on the reference path")
        return var

    def step(self, model_output, timestep, sample, noise=None, generator=None):
        """Returns (x_{t-1}, x0_hat). `noise` injects z for parity; otherwise drawn on CPU."""
        t = int(timestep)
        acp_t = self.alphas_cumprod[t]
        acp_prev = self.alphas_cumprod[t - 1] if t > 0 else self.one
        beta_prod_t = 1 - acp_t
        beta_prod_prev = 1 - acp_prev
        if self.prediction_type == "epsilon":
            x0 = (sample - beta_prod_t ** 0.5 * model_output) / acp_t ** 0.5
        elif self.prediction_type == "sample":
            x0 = model_output
        elif self.prediction_type == "v_prediction":
            x0 = (acp_t ** 0.5) * sample - (beta_prod_t ** 0.5) * model_output
        else:
            raise ValueError(self.prediction_type)
        if self.clip_sample:
            x0 = torch.clamp(x0, -1, 1)
        c0 = (acp_prev ** 0.5 * self.betas[t]) / beta_prod_t
        ct = self.alphas[t] ** 0.5 * beta_prod_prev / beta_prod_t
        prev = c0 * x0 + ct * sample
        if t > 0:
            if noise is None:
                noise = torch.randn(model_output.size(), dtype=model_output.dtype, generator=generator)
            prev = prev + (self._variance(t) ** 0.5) * noise
        return prev, x0


class OracleDiffusionInferer:
    def __init__(self, scheduler):
        self.scheduler = scheduler

    def __call__(self, inputs, diffusion_model, noise, timesteps, condition=None):
        noisy = self.scheduler.add_noise(original_samples=inputs, noise=noise, timesteps=timesteps)
        return diffusion_model(noisy, timesteps=timesteps, context=condition)

    @torch.no_grad()
    def sample(self, input_noise, diffusion_model, scheduler=None, conditioning=None, step_noises=None):
        scheduler = scheduler or self.scheduler
        image = input_noise
        for i, t in enumerate(scheduler.timesteps):
            out = diffusion_model(image, timesteps=torch.Tensor((t,)), context=conditioning)
            z = None if step_noises is None else step_noises[i]
            image, _ = scheduler.step(out, t, image, noise=z)
        return image


class OracleLatentDiffusionInferer(OracleDiffusionInferer):
    def __init__(self, scheduler, scale_factor: float = 1.0):
        super().__init__(scheduler)
        self.scale_factor = scale_factor

    def __call__(self, inputs, autoencoder_model, diffusion_model, noise, timesteps, condition=None):
        with torch.no_grad():
            latent = autoencoder_model.encode_stage_2_inputs(inputs) * self.scale_factor
        return super().__call__(latent, diffusion_model, noise, timesteps, condition)

    @torch.no_grad()
    def sample(self, input_noise, autoencoder_model, diffusion_model, scheduler=None, conditioning=None,
               step_noises=None):
        latent = super().sample(input_noise, diffusion_model, scheduler, conditioning, step_noises)
        return autoencoder_model.decode_stage_2_outputs(latent / self.scale_factor)
