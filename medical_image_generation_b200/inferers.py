"""DiffusionInferer / LatentDiffusionInferer with monai-generative's interface (the external dependency used at
train_ddpm.py:191,243 and train_ldm.py:112,157,362), driving the B200 modules and the fused scheduler kernels."""
from __future__ import annotations

import torch

from . import ops


class DiffusionInferer:
    def __init__(self, scheduler) -> None:
        self.scheduler = scheduler

    def __call__(self, inputs, diffusion_model, noise, timesteps, condition=None, mode: str = "crossattn"):
        if mode not in ("crossattn", "concat"):
            raise NotImplementedError(f"{mode} condition is not supported")
        noisy = self.scheduler.add_noise(original_samples=inputs, noise=noise, timesteps=timesteps)
        if mode == "concat":
            noisy = ops.cat_channels(noisy, condition)
            condition = None
        return diffusion_model(x=noisy, timesteps=timesteps, context=condition)

    @torch.no_grad()
    def sample(self, input_noise, diffusion_model, scheduler=None, save_intermediates: bool = False,
               intermediate_steps: int = 100, conditioning=None, mode: str = "crossattn", verbose: bool = True,
               step_noises=None, cuda_graph: bool = False):
        """Reverse process over `scheduler.timesteps`; the model is called with `torch.Tensor((t,))` like upstream.
        `step_noises` (optional list) injects the per-step z for parity tests. `cuda_graph=True` (extra, optional)
        captures the model forward once and replays it every step: the ~300 kernel launches of a reverse step stop
        costing host time (the fused scheduler step stays eager: its coefficients change every step)."""
        if mode not in ("crossattn", "concat"):
            raise NotImplementedError(f"{mode} condition is not supported")
        scheduler = scheduler or self.scheduler
        image = input_noise
        intermediates = []

        def forward(img, tt):
            if mode == "concat":
                return diffusion_model(ops.cat_channels(img, conditioning), timesteps=tt, context=None)
            return diffusion_model(img, timesteps=tt, context=conditioning)

        graph = None
        if cuda_graph and input_noise.is_cuda and len(scheduler.timesteps) > 2:
            static_img = input_noise.clone()
            static_t = torch.zeros(1, dtype=torch.float32, device=input_noise.device)
            forward(static_img, static_t)                      # warm-up: workspaces, filter shadows, tensor maps
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                static_out = forward(static_img, static_t)
        for i, t in enumerate(scheduler.timesteps):
            if graph is not None:
                static_img.copy_(image)
                static_t.fill_(float(t))
                graph.replay()
                out = static_out
            else:
                out = forward(image, torch.Tensor((t,)).to(input_noise.device))
            z = None if step_noises is None else step_noises[i]
            image, _ = scheduler.step(out, t, image, noise=z)
            if save_intermediates and t % intermediate_steps == 0:
                intermediates.append(image)
        return (image, intermediates) if save_intermediates else image


class LatentDiffusionInferer(DiffusionInferer):
    def __init__(self, scheduler, scale_factor: float = 1.0) -> None:
        super().__init__(scheduler=scheduler)
        self.scale_factor = scale_factor

    def __call__(self, inputs, autoencoder_model, diffusion_model, noise, timesteps, condition=None,
                 mode: str = "crossattn"):
        with torch.no_grad():
            latent = ops.scale(autoencoder_model.encode_stage_2_inputs(inputs), float(self.scale_factor))
        return super().__call__(inputs=latent, diffusion_model=diffusion_model, noise=noise, timesteps=timesteps,
                                condition=condition, mode=mode)

    @torch.no_grad()
    def sample(self, input_noise, autoencoder_model, diffusion_model, scheduler=None, save_intermediates: bool = False,
               intermediate_steps: int = 100, conditioning=None, mode: str = "crossattn", verbose: bool = True,
               step_noises=None, cuda_graph: bool = False):
        out = super().sample(input_noise, diffusion_model, scheduler, save_intermediates, intermediate_steps,
                             conditioning, mode, verbose, step_noises, cuda_graph)
        latent, inter = out if save_intermediates else (out, None)
        image = autoencoder_model.decode_stage_2_outputs(ops.scale(latent, 1.0 / float(self.scale_factor)))
        if save_intermediates:
            return image, [autoencoder_model.decode_stage_2_outputs(ops.scale(l, 1.0 / float(self.scale_factor)))
                           for l in inter]
        return image
