"""Data-parallel training engine for the LDM / DDPM step (train_ldm.py:143-183, train_ddpm.py:183-201 semantics).

B200-first layout of the optimiser state: ONE flat fp32 buffer each for master parameters, gradients, Adam m and v
and ONE flat bf16 "shadow" of the parameters. Every nn.Parameter becomes a strided view into the master buffer
(filters keep their channels-last [Cout][taps][Cin] strides), so:
  * the weight-gradient kernels accumulate straight into the flat gradient buffer (no per-parameter zero/accumulate),
  * gradient clipping is one sum-of-squares launch, AdamW is one fused launch that also refreshes the bf16 shadow the
    tensor-core kernels read (no per-step fp32->bf16 filter casts),
  * gradient buckets for the NCCL all-reduce are contiguous slices -- launched from the backward pass as soon as every
    gradient of a bucket has been produced, on torch.distributed's communication stream (overlap with backward).
Parameters that never receive a gradient (`proj_attn`) are placed at the tail of the buffers and are
skipped by the optimiser exactly like torch.optim.AdamW skips `grad is None`.
"""
from __future__ import annotations

import ctypes as C
from typing import Iterable, Optional

import torch
import torch.distributed as dist

from . import ops
from ._lib import call


def _dist_on() -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


class GradBuckets:
    """Contiguous gradient buckets over ONE flat buffer, all-reduced (mean) as soon as every member gradient has
    been produced. Device-agnostic host logic: NCCL over NVLink on the GPUs, gloo in the CPU tests.

    `sizes[i]` is the (padded) element count of member i in buffer order; members are referred to by index."""

    def __init__(self, flat: torch.Tensor, sizes, cap_elems: int, group=None):
        self.flat, self.group = flat, group
        self.world = dist.get_world_size(group)
        self.avg_native = dist.get_backend(group) == "nccl"   # gloo has no ReduceOp.AVG
        self.buckets, self.bucket_of = [], []
        start = count = n = 0
        for sz in sizes:
            self.bucket_of.append(len(self.buckets))
            count += int(sz)
            n += 1
            if count >= cap_elems:
                self.buckets.append(dict(lo=start, hi=start + count, n=n, pending=n, work=None))
                start, count, n = start + count, 0, 0
        if n:
            self.buckets.append(dict(lo=start, hi=start + count, n=n, pending=n, work=None))
        self.seen = set()

    def _launch(self, bk) -> None:
        op = dist.ReduceOp.AVG if self.avg_native else dist.ReduceOp.SUM
        bk["work"] = dist.all_reduce(self.flat[bk["lo"]:bk["hi"]], op=op, group=self.group, async_op=True)

    def ready(self, index: int) -> None:
        """Member `index` has its final gradient in the flat buffer."""
        if index in self.seen:
            return
        self.seen.add(index)
        bk = self.buckets[self.bucket_of[index]]
        bk["pending"] -= 1
        if bk["pending"] == 0:
            self._launch(bk)

    def reset(self) -> None:
        self.seen.clear()
        for bk in self.buckets:
            bk["pending"], bk["work"] = bk["n"], None

    def finish(self) -> None:
        """Reduce the buckets that did not fire during backward (a member unused this step), then wait for all."""
        for bk in self.buckets:
            if bk["work"] is None:
                self._launch(bk)
        for bk in self.buckets:
            bk["work"].wait()
            if not self.avg_native:
                self.flat[bk["lo"]:bk["hi"]].div_(self.world)


class FlatAdamW:
    """AdamW + global-norm clipping over flat buffers; numerics follow torch.optim.AdamW / clip_grad_norm_."""

    def __init__(self, module: torch.nn.Module, lr: float = 2e-5, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 1e-2, max_grad_norm: Optional[float] = 1.0, bucket_mb: float = 64.0,
                 unused: Iterable[str] = ("proj_attn",)):
        named = [(n, p) for n, p in module.named_parameters() if p.requires_grad]
        if not named:
            raise ValueError("module has no trainable parameters")
        self._torch_order = [p for _, p in named]   # torch.optim indexes parameters in registration order
        dev = named[0][1].device
        if dev.type != "cuda":
            raise RuntimeError("FlatAdamW needs the module on a CUDA device")
        is_unused = lambda n: any(tag in n for tag in unused)  # noqa: E731
        # backward produces gradients roughly in reverse registration order: lay the buffer out in that order so a
        # bucket (contiguous slice) completes early and can be reduced while the rest of backward still runs
        used = [(n, p) for n, p in reversed(named) if not is_unused(n)]
        tail = [(n, p) for n, p in named if is_unused(n)]
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.max_grad_norm = max_grad_norm
        self.step_count = 0
        # every parameter starts on a 64-element boundary (256 B fp32 / 128 B bf16): TMA and the 16-byte vector paths
        # need aligned bases. Padding elements stay exactly zero (zero grad, zero master) through AdamW.
        pad = lambda k: (k + 63) // 64 * 64  # noqa: E731
        self.used_numel = sum(pad(p.numel()) for _, p in used)
        total = self.used_numel + sum(pad(p.numel()) for _, p in tail)
        self.master = torch.zeros(total, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(total, dtype=torch.float32, device=dev)
        self.m = torch.zeros(self.used_numel, dtype=torch.float32, device=dev)
        self.v = torch.zeros(self.used_numel, dtype=torch.float32, device=dev)
        self.shadow = torch.empty(total, dtype=torch.bfloat16, device=dev)
        self.sumsq = torch.zeros(1, dtype=torch.float32, device=dev)
        self._partials = torch.zeros(2048, dtype=torch.float32, device=dev)
        self.step_dev = torch.zeros(1, dtype=torch.int32, device=dev)
        self.params = []
        off = 0
        for n, p in used + tail:
            k = p.numel()
            with torch.no_grad():
                view = self.master[off:off + k].as_strided(p.shape, p.stride())
                view.copy_(p)
                p.data = view
            p.main_grad = self.grad[off:off + k].as_strided(p.shape, p.stride())
            p._mig_shadow = self.shadow[off:off + k].as_strided(p.shape, p.stride())
            p._mig_shadow_version = p._version   # ops._filter_for re-casts the slot when the master changed in place
            p._mig_slot = (off, k)
            self.params.append((n, p))
            off += pad(k)
        call("mig_cast", 0, 1, ops._ptr(self.master), ops._ptr(self.shadow), total, ops._stream())
        # ---- gradient buckets (contiguous slices of the used region) ----
        self.buckets = None
        self._index_of = {}
        self._sync_enabled = True
        if _dist_on():
            self.buckets = GradBuckets(self.grad, [pad(p.numel()) for _, p in used], int(bucket_mb * (1 << 20) / 4))
            self._index_of = {id(p): i for i, (_, p) in enumerate(used)}
        # every owned parameter carries the callback of ITS optimiser (several FlatAdamW instances can coexist)
        for _, p in self.params:
            p._mig_grad_ready = self._grad_ready

    # called from the backward kernels' wrappers once a parameter's gradient is complete in `main_grad`
    def _grad_ready(self, p) -> None:
        if self.buckets is None or not self._sync_enabled:   # accumulation: only the last micro-step reduces
            return
        i = self._index_of.get(id(p))
        if i is not None:
            self.buckets.ready(i)

    def set_grad_sync(self, enabled: bool) -> None:
        """Gradient accumulation (train_ldm.py:173): micro-steps before the last one add into the flat gradient
        buffer without firing the bucket all-reduce; the last micro-step's backward reduces the accumulated sums."""
        self._sync_enabled = bool(enabled)

    def zero_grad(self) -> None:
        self.grad.zero_()
        if self.buckets is not None:
            self.buckets.reset()

    def finish_grad_sync(self) -> None:
        self.buckets.finish()

    def step(self) -> None:
        # gradients that arrived through plain autograd (.grad) -- e.g. nn.Embedding -- join the flat buffer here
        for _, p in self.params:
            if p.grad is not None:
                p.main_grad.add_(p.grad)
                p.grad = None
        if self.buckets is not None:
            self.finish_grad_sync()
        self.step_count += 1          # host mirror; the kernel reads the device counter (valid under graph replay)
        self.step_dev.add_(1)
        st = ops._stream()
        sumsq_ptr = None
        max_norm = 0.0
        if self.max_grad_norm:
            call("mig_sumsq", ops._ptr(self.grad), ops._ptr(self.sumsq), ops._ptr(self._partials), self.used_numel, st)
            sumsq_ptr, max_norm = ops._ptr(self.sumsq), float(self.max_grad_norm)
        call("mig_adamw_step", ops._ptr(self.master), ops._ptr(self.grad), ops._ptr(self.m), ops._ptr(self.v),
             self.used_numel, float(self.lr), float(self.betas[0]), float(self.betas[1]), float(self.eps),
             float(self.weight_decay), int(self.step_count), sumsq_ptr, max_norm, ops._ptr(self.shadow),
             ops._ptr(self.step_dev), st)

    def refresh_shadow(self) -> None:
        """Re-derive the whole bf16 shadow from the fp32 master (one cast launch). Needed only after writing to
        `self.master` directly; in-place changes through the nn.Parameters (load_state_dict, copy_, initialize) are
        detected per parameter by ops._filter_for via the autograd version counter."""
        call("mig_cast", 0, 1, ops._ptr(self.master), ops._ptr(self.shadow), self.master.numel(), ops._stream())
        for _, p in self.params:
            p._mig_shadow_version = p._version

    def grad_norm(self) -> torch.Tensor:
        """Global gradient norm of the last step() (device scalar; no sync)."""
        return self.sumsq.sqrt()

    def state_dict(self) -> dict:
        return dict(step=self.step_count, m=self.m, v=self.v, lr=self.lr)

    # ---- checkpoint interchange with the reference: ckpt['optimizer_state_dict'] (train_ldm.py:472-477, 496-499) ----
    def torch_state_dict(self) -> dict:
        """The optimiser state in torch.optim.AdamW's own state_dict() layout (parameters indexed in registration
        order, per-parameter `step` / `exp_avg` / `exp_avg_sq`; parameters that never received a gradient have no
        entry, as in torch), so a checkpoint written here resumes in the reference trainer and vice versa."""
        state = {}
        for i, p in enumerate(self._torch_order):
            off, k = p._mig_slot
            if off >= self.used_numel or self.step_count == 0:
                continue
            view = lambda buf: buf[off:off + k].as_strided(p.shape, p.stride()).detach().clone()  # noqa: E731
            state[i] = {"step": torch.tensor(float(self.step_count)), "exp_avg": view(self.m), "exp_avg_sq": view(self.v)}
        group = dict(lr=self.lr, betas=tuple(self.betas), eps=self.eps, weight_decay=self.weight_decay, amsgrad=False,
                     maximize=False, foreach=None, capturable=False, differentiable=False, fused=None,
                     params=list(range(len(self._torch_order))))
        return {"state": state, "param_groups": [group]}

    def load_torch_state_dict(self, sd: dict) -> None:
        """Inverse of torch_state_dict(): accepts torch.optim.AdamW / Adam state dicts of the same module."""
        group = sd["param_groups"][0]
        if len(group["params"]) != len(self._torch_order):
            raise ValueError(f"optimizer state has {len(group['params'])} parameters, the module has {len(self._torch_order)}")
        self.lr, self.betas, self.eps = float(group["lr"]), tuple(group["betas"]), float(group["eps"])
        self.weight_decay = float(group.get("weight_decay", 0.0))
        self.m.zero_()
        self.v.zero_()
        step = 0
        for j, idx in enumerate(group["params"]):
            st = sd["state"].get(idx)
            if st is None:
                continue
            p = self._torch_order[j]
            off, k = p._mig_slot
            if off >= self.used_numel:
                continue
            with torch.no_grad():
                self.m[off:off + k].as_strided(p.shape, p.stride()).copy_(st["exp_avg"])
                self.v[off:off + k].as_strided(p.shape, p.stride()).copy_(st["exp_avg_sq"])
            step = max(step, int(float(st["step"])))
        self.step_count = step
        self.step_dev.fill_(step)

    def close(self) -> None:
        for _, p in self.params:
            if getattr(p, "_mig_grad_ready", None) == self._grad_ready:
                p._mig_grad_ready = None


class LDMTrainer:
    """One optimiser micro-step of train_ldm.LDM.train_one_epoch (train_ldm.py:143-183) on latents already encoded
    and scaled (the frozen autoencoder's `encode_stage_2_inputs(images) * scale_factor` is run by the caller, under
    no_grad, exactly as the reference does), data-parallel across the process group when one is initialised."""

    def __init__(self, unet, scheduler, lr: float = 2e-5, grad_clip_max_norm: Optional[float] = 1.0,
                 weight_decay: float = 1e-2, bucket_mb: float = 64.0, cuda_graph: bool = False,
                 graph_warmup_steps: int = 3, grad_accumulate_step: int = 1):
        self.unet, self.scheduler = unet, scheduler
        # config['grad_accumulate_step'] (train_ldm.py:173): the optimiser steps every k-th call of step(); gradients
        # of the micro-steps are SUMMED (the reference does not rescale the loss)
        self.grad_accumulate_step = max(1, int(grad_accumulate_step))
        self._micro = 0
        self.opt = FlatAdamW(unet, lr=lr, weight_decay=weight_decay, max_grad_norm=grad_clip_max_norm,
                             bucket_mb=bucket_mb)
        if _dist_on():  # identical replicas: broadcast rank 0's parameters (flat: one collective)
            dist.broadcast(self.opt.master, src=0)
            self.opt.refresh_shadow()
        # CUDA graph of the whole step (add_noise -> U-Net fwd -> MSE -> bwd -> all-reduce -> clip -> AdamW): the
        # step is ~1300 kernel launches, so replaying one graph removes the launch gaps. The first
        # `graph_warmup_steps` calls run eagerly (they are real optimiser steps), the next call is captured.
        self.cuda_graph = cuda_graph and self.grad_accumulate_step == 1
        self._eager_calls = 0
        self._warm = graph_warmup_steps
        self._graph = None
        self._static_x = self._static_loss = None
        self._graph_key = None

    def step(self, latents_scaled: torch.Tensor, noise: Optional[torch.Tensor] = None,
             timesteps: Optional[torch.Tensor] = None) -> torch.Tensor:
        if self.cuda_graph and noise is None and timesteps is None:
            return self._graph_step(latents_scaled)
        return self._eager_step(latents_scaled, noise, timesteps)

    def _graph_step(self, x0: torch.Tensor) -> torch.Tensor:
        key = (tuple(x0.shape), x0.dtype, float(self.opt.lr))
        if self._graph is not None and key != self._graph_key:   # shape or lr changed: capture again
            self._graph, self._eager_calls = None, 0
        if self._graph is None:
            if self._eager_calls < self._warm:
                self._eager_calls += 1
                return self._eager_step(x0, None, None)
            self._static_x = x0.clone()
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            saved = (self.opt.step_count, self._micro)   # host state the captured (never executed) step advances
            try:
                with torch.cuda.graph(graph):
                    self._static_loss = self._eager_step(self._static_x, None, None)
            except Exception as e:  # noqa: BLE001 -- fall back loudly, never silently change numerics
                self.cuda_graph = False
                self.opt.step_count, self._micro = saved
                if self.opt.buckets is not None:
                    self.opt.buckets.reset()
                import warnings
                warnings.warn(f"CUDA-graph capture of the training step failed ({e}); continuing eagerly")
                torch.cuda.synchronize()
                return self._eager_step(x0, None, None)
            self._graph, self._graph_key = graph, key
            # capture does not execute: the captured step's host-side counter advanced once, the device one did not
        else:
            self._static_x.copy_(x0, non_blocking=True)
            self.opt.step_count += 1
        self._graph.replay()
        return self._static_loss

    def _eager_step(self, latents_scaled: torch.Tensor, noise: Optional[torch.Tensor] = None,
                    timesteps: Optional[torch.Tensor] = None) -> torch.Tensor:
        s = self.scheduler
        x0 = latents_scaled
        if timesteps is None:  # train_ldm.py:145
            timesteps = torch.randint(0, s.num_train_timesteps, (x0.shape[0],), device=x0.device).long()
        if noise is None:      # train_ldm.py:159
            noise = torch.randn_like(x0)
        last = self._micro + 1 >= self.grad_accumulate_step
        if self._micro == 0:
            self.opt.zero_grad()
        self.opt.set_grad_sync(last)
        noisy = s.add_noise(original_samples=x0, noise=noise, timesteps=timesteps)
        pred = self.unet(x=noisy, timesteps=timesteps)
        target = s.get_velocity(x0, noise, timesteps) if s.prediction_type == "v_prediction" else noise
        loss = ops.mse_loss(pred, target)
        loss.backward()
        if last:
            self.opt.step()
            self._micro = 0
        else:
            self._micro += 1
        return loss

    def flush(self) -> None:
        """Optimiser step on a partial accumulation group (the reference also steps on the last batch of an epoch,
        train_ldm.py:173: `or (step + 1) == len(train_loader)`)."""
        if self._micro > 0:
            self.opt.set_grad_sync(True)
            self.opt.step()
            self._micro = 0


class AETrainer:
    """Generator step of train_autoencoder.AutoEncoder.train_generator_step (train_autoencoder.py:399-436) restricted
    to the terms that need no downloaded networks: L1 reconstruction + kl_weight * KL (train_autoencoder.py:67-72,
    412-414). Perceptual and adversarial terms are out of scope. Data-parallel like LDMTrainer."""

    def __init__(self, autoencoder, lr: float = 5e-5, kl_weight: float = 1e-7,
                 grad_clip_max_norm: Optional[float] = 1.0, bucket_mb: float = 64.0, grad_accumulate_step: int = 1):
        self.ae, self.kl_weight = autoencoder, kl_weight
        self.grad_accumulate_step = max(1, int(grad_accumulate_step))   # train_autoencoder.py:427
        self._micro = 0
        # the reference uses torch.optim.Adam (no weight decay) for the generator (train_autoencoder.py:470)
        self.opt = FlatAdamW(autoencoder, lr=lr, weight_decay=0.0, max_grad_norm=grad_clip_max_norm,
                             bucket_mb=bucket_mb, unused=("proj_attn",))
        if _dist_on():
            dist.broadcast(self.opt.master, src=0)
            self.opt.refresh_shadow()

    def step(self, images: torch.Tensor, eps: Optional[torch.Tensor] = None) -> torch.Tensor:
        last = self._micro + 1 >= self.grad_accumulate_step
        if self._micro == 0:
            self.opt.zero_grad()
        self.opt.set_grad_sync(last)
        z_mu, z_sigma = self.ae.encode(images)
        z = ops.vae_reparam(z_mu, z_sigma, eps) if eps is not None else self.ae.sampling(z_mu, z_sigma)
        recon = self.ae.decode(z)
        loss = ops.l1_loss(recon, images) + self.kl_weight * ops.kl_loss(z_mu, z_sigma)
        loss.backward()
        if last:
            self.opt.step()
            self._micro = 0
        else:
            self._micro += 1
        return loss

    def flush(self) -> None:
        """Optimiser step on a trailing partial accumulation group (train_autoencoder.py:427)."""
        if self._micro > 0:
            self.opt.set_grad_sync(True)
            self.opt.step()
            self._micro = 0


@torch.no_grad()
def sample_volumes(unet, scheduler, shape, num_volumes: int, base_seed: int = 42, autoencoder=None,
                   scale_factor: float = 1.0, num_inference_steps: Optional[int] = None, noise_mode: str = "device"):
    """Sampling sharder: volume v goes to rank v % world, seeded base_seed + v, full reverse process
    (train_ldm.py:332-366 / train_ddpm.py:238-246) on that rank; NO communication. Returns {volume index: tensor} for
    the volumes of this rank. `shape` is one volume's (C, *spatial)."""
    from .inferers import DiffusionInferer, LatentDiffusionInferer
    rank = dist.get_rank() if (dist.is_available() and dist.is_initialized()) else 0
    world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
    dev = next(unet.parameters()).device
    scheduler.noise_mode = noise_mode
    scheduler.set_timesteps(num_inference_steps or scheduler.num_train_timesteps)
    unet.eval()
    out = {}
    for v in range(rank, num_volumes, world):
        with torch.random.fork_rng(devices=[dev]):
            torch.manual_seed(base_seed + v)                                   # per-volume seed
            noise = torch.randn(1, *shape).to(dev)                             # CPU draw like train_ldm.py:343-349
            if autoencoder is not None:
                inf = LatentDiffusionInferer(scheduler, scale_factor=scale_factor)
                out[v] = inf.sample(noise, autoencoder_model=autoencoder, diffusion_model=unet, scheduler=scheduler,
                                    verbose=False)
            else:
                out[v] = DiffusionInferer(scheduler).sample(noise, unet, scheduler, verbose=False)
    return out
