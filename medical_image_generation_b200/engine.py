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

import os

import ctypes as C
from typing import Iterable, Optional

import torch
import torch.distributed as dist

from . import ops
from ._lib import call


_COALESCE_GATHER = os.environ.get("MIG_COALESCE_GATHER", "1") != "0"   # A/B switch, see ShardedBuckets.gather


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


class ShardedBuckets:
    """Gradient exchange of the SHARDED optimiser (ZeRO-1 style; SURVEY.md section 8e asks for one exchange step -- this is
    the same step with half the wire bytes and 1/world of the optimiser work per rank).

    The flat buffer is [ sharded region | replicated region | ... ]:
      * sharded region = the bf16-consumed parameters (conv filters, attention projections): `nb` UNIFORM buckets of
        `bucket` elements; bucket b is reduce-scattered (mean) as soon as every parameter overlapping it has its
        gradient, so rank r ends up with the reduced slice [b*bucket + r*piece, +piece) of every bucket (piece =
        bucket / world). A parameter may span several buckets.
      * replicated region = the few parameters the kernels read in fp32 (biases, norm gains, the time-embedding MLP):
        ONE all-reduce; every rank updates them redundantly, so no parameter gather is needed for them.
    After the optimiser step `gather(buf)` all-gathers the owned slices of `buf` (the bf16 shadow every step; the fp32
    master / Adam moments only for checkpoints). gloo (CPU tests) has neither AVG nor reduce_scatter_tensor: it falls back
    to all_reduce + divide, which leaves the same values in the owned slices."""

    def __init__(self, flat: torch.Tensor, spans, shard_numel: int, bucket: int, repl_lo: int, repl_hi: int,
                 repl_members, group=None):
        self.flat, self.group = flat, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.nccl = dist.get_backend(group) == "nccl"
        if bucket % self.world or shard_numel % bucket:
            raise ValueError("bucket size must divide the sharded region and be a multiple of the world size")
        self.bucket, self.piece, self.nb = bucket, bucket // self.world, shard_numel // bucket
        self.shard_numel = shard_numel
        self.buckets = [dict(lo=b * bucket, hi=(b + 1) * bucket, kind="rs", n=0, pending=0, work=None)
                        for b in range(self.nb)]
        self.member_buckets = []                      # member index -> bucket indices it overlaps
        for lo, hi in spans:                          # sharded members, buffer order
            bs = list(range(lo // bucket, (hi - 1) // bucket + 1)) if hi > lo else []
            for b in bs:
                self.buckets[b]["n"] += 1
            self.member_buckets.append(bs)
        self.repl = None
        if repl_hi > repl_lo:
            self.repl = dict(lo=repl_lo, hi=repl_hi, kind="ar", n=len(repl_members), pending=0, work=None)
            self.buckets.append(self.repl)
            for _ in repl_members:
                self.member_buckets.append([len(self.buckets) - 1])
        self.seen = set()
        self.reset()

    def _launch(self, bk) -> None:
        buf = self.flat[bk["lo"]:bk["hi"]]
        if bk["kind"] == "rs" and self.nccl:
            mine = buf[self.rank * self.piece:(self.rank + 1) * self.piece]     # in place: output = own slice of input
            bk["work"] = dist.reduce_scatter_tensor(mine, buf, op=dist.ReduceOp.AVG, group=self.group, async_op=True)
        else:
            op = dist.ReduceOp.AVG if self.nccl else dist.ReduceOp.SUM
            bk["work"] = dist.all_reduce(buf, op=op, group=self.group, async_op=True)

    def ready(self, index: int) -> None:
        if index in self.seen:
            return
        self.seen.add(index)
        for b in self.member_buckets[index]:
            bk = self.buckets[b]
            bk["pending"] -= 1
            if bk["pending"] == 0:
                self._launch(bk)

    def reset(self) -> None:
        self.seen.clear()
        for bk in self.buckets:
            bk["pending"], bk["work"] = bk["n"], None

    def finish(self) -> None:
        for bk in self.buckets:
            if bk["work"] is None:
                self._launch(bk)
        for bk in self.buckets:
            bk["work"].wait()
            if not self.nccl:
                self.flat[bk["lo"]:bk["hi"]].div_(self.world)

    def owned(self, buf: torch.Tensor, b: int) -> torch.Tensor:
        lo = b * self.bucket + self.rank * self.piece
        return buf[lo:lo + self.piece]

    def gather(self, buf: torch.Tensor) -> None:
        """All-gather the owned slices of `buf` (same layout as the gradient buffer) in place, bucket by bucket."""
        if self.nccl and self.nb > 1 and _COALESCE_GATHER:
            # ONE NCCL group (one kernel) for all buckets: issued one by one, the gathers of the 26 LDM buckets were 26
            # launch + handshake latencies in a row at the very end of the step, where nothing overlaps them
            # (tools/dp_timeline.py on 2 GPUs: 26 x 52 us = 1.4 ms for 0.44 GB)
            from torch.distributed.distributed_c10d import _coalescing_manager
            with _coalescing_manager(group=self.group, device=buf.device, async_ops=True) as cm:
                for b in range(self.nb):
                    dist.all_gather_into_tensor(buf[b * self.bucket:(b + 1) * self.bucket], self.owned(buf, b),
                                                group=self.group)
            cm.wait()
            return
        works = []
        for b in range(self.nb):
            whole = buf[b * self.bucket:(b + 1) * self.bucket]
            if self.nccl:
                works.append(dist.all_gather_into_tensor(whole, self.owned(buf, b), group=self.group, async_op=True))
            else:
                parts = [torch.empty_like(self.owned(buf, b)) for _ in range(self.world)]
                dist.all_gather(parts, self.owned(buf, b).clone(), group=self.group)
                for r, part in enumerate(parts):
                    whole[r * self.piece:(r + 1) * self.piece].copy_(part)
        for w in works:
            w.wait()


class FlatAdamW:
    """AdamW + global-norm clipping over flat buffers; numerics follow torch.optim.AdamW / clip_grad_norm_."""

    FP32_CONSUMED = ("time_embed", "time_emb_proj", "class_embedding")   # Linear weights the kernels read in fp32

    def __init__(self, module: torch.nn.Module, lr: float = 2e-5, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 1e-2, max_grad_norm: Optional[float] = 1.0, bucket_mb: float = 64.0,
                 unused: Iterable[str] = ("proj_attn",), shard: Optional[bool] = None):
        named = [(n, p) for n, p in module.named_parameters() if p.requires_grad]
        if not named:
            raise ValueError("module has no trainable parameters")
        self._torch_order = [p for _, p in named]   # torch.optim indexes parameters in registration order
        dev = named[0][1].device
        if dev.type != "cuda":
            raise RuntimeError("FlatAdamW needs the module on a CUDA device")
        is_unused = lambda n: any(tag in n for tag in unused)  # noqa: E731
        world = dist.get_world_size() if _dist_on() else 1
        # Sharded update (ZeRO-1 style) when data parallel: only meaningful in bf16 mode, where every large parameter
        # is consumed through its bf16 shadow (the fp32 parity mode reads the masters themselves on every rank).
        bf16_mode = getattr(module, "compute_dtype", torch.bfloat16) == torch.bfloat16
        self.sharded = world > 1 and bf16_mode and (True if shard is None else bool(shard))
        # backward produces gradients roughly in reverse registration order: lay the buffer out in that order so a
        # bucket (contiguous slice) completes early and can be reduced while the rest of backward still runs
        used = [(n, p) for n, p in reversed(named) if not is_unused(n)]
        tail = [(n, p) for n, p in named if is_unused(n)]
        # fp32-consumed parameters (1-D: biases / norm gains; the time-embedding MLP) form the replicated region
        fp32_used = lambda n, p: p.ndim <= 1 or any(tag in n for tag in self.FP32_CONSUMED)  # noqa: E731
        # (the split is kept without sharding too: it puts to_q / to_k / to_v weights -- and their biases -- next to each
        #  other, which is what lets an attention block run ONE fused QKV projection, ops.fuse_linears)
        shard_part = [(n, p) for n, p in used if not fp32_used(n, p)]
        repl_part = [(n, p) for n, p in used if fp32_used(n, p)]
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.max_grad_norm = max_grad_norm
        self.step_count = 0
        # every parameter starts on a 64-element boundary (256 B fp32 / 128 B bf16): TMA and the 16-byte vector paths
        # need aligned bases. Padding elements stay exactly zero (zero grad, zero master) through AdamW.
        pad = lambda k: (k + 63) // 64 * 64  # noqa: E731
        shard_raw = sum(pad(p.numel()) for _, p in shard_part)
        self.bucket_elems = 0
        self.shard_numel = shard_raw
        if self.sharded and shard_raw:
            unit = world * 64
            self.bucket_elems = max(unit, (int(bucket_mb * (1 << 20) / 4) + unit - 1) // unit * unit)
            self.shard_numel = (shard_raw + self.bucket_elems - 1) // self.bucket_elems * self.bucket_elems
        repl_numel = sum(pad(p.numel()) for _, p in repl_part)
        self.repl_lo = self.shard_numel
        self.used_numel = self.shard_numel + repl_numel
        total = self.used_numel + sum(pad(p.numel()) for _, p in tail)
        self.master = torch.zeros(total, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(total, dtype=torch.float32, device=dev)
        self.m = torch.zeros(self.used_numel, dtype=torch.float32, device=dev)
        self.v = torch.zeros(self.used_numel, dtype=torch.float32, device=dev)
        self.shadow = torch.empty(total, dtype=torch.bfloat16, device=dev)
        self.sumsq = torch.zeros(1, dtype=torch.float32, device=dev)
        self._ss = torch.zeros(2, dtype=torch.float32, device=dev)     # sharded: [own slices, replicated region]
        self._partials = torch.zeros(2048, dtype=torch.float32, device=dev)
        self.step_dev = torch.zeros(1, dtype=torch.int32, device=dev)
        self.params = []

        def place(group, off):
            spans = []
            for n, p in group:
                k = p.numel()
                with torch.no_grad():
                    view = self.master[off:off + k].as_strided(p.shape, p.stride())
                    view.copy_(p)
                    p.data = view
                p.main_grad = self.grad[off:off + k].as_strided(p.shape, p.stride())
                p._mig_shadow = self.shadow[off:off + k].as_strided(p.shape, p.stride())
                p._mig_shadow_version = p._version   # ops._filter_for re-casts the slot when the master changed in place
                p._mig_slot = (off, k)
                p._mig_flat = self
                self.params.append((n, p))
                spans.append((off, off + pad(k)))
                off += pad(k)
            return spans, off

        shard_spans, _ = place(shard_part, 0)
        _, off = place(repl_part, self.repl_lo)
        place(tail, off)
        call("mig_cast", 0, 1, ops._ptr(self.master), ops._ptr(self.shadow), total, ops._stream())
        # ---- gradient buckets (contiguous slices of the used region) ----
        self.buckets = None
        self._index_of = {}
        self._sync_enabled = True
        if _dist_on():
            if self.sharded:
                self.buckets = ShardedBuckets(self.grad, shard_spans, self.shard_numel, self.bucket_elems or world * 64,
                                              self.repl_lo, self.used_numel, repl_part)
                self._index_of = {id(p): i for i, (_, p) in enumerate(shard_part + repl_part)}
            else:
                order = shard_part + repl_part
                self.buckets = GradBuckets(self.grad, [pad(p.numel()) for _, p in order], int(bucket_mb * (1 << 20) / 4))
                self._index_of = {id(p): i for i, (_, p) in enumerate(order)}
        # every owned parameter carries the callback of ITS optimiser (several FlatAdamW instances can coexist)
        for _, p in self.params:
            p._mig_grad_ready = self._grad_ready
        self._seen = set()          # ids of parameters that have delivered a gradient at least once
        self._checked_unused = False

    # called from the backward kernels' wrappers once a parameter's gradient is complete in `main_grad`
    def _grad_ready(self, p) -> None:
        self._seen.add(id(p))
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
                self._seen.add(id(p))
        if not self._checked_unused:
            # torch.optim.AdamW skips parameters whose grad is None (no weight decay, no moments). The flat kernel updates
            # the whole used region, so a trainable parameter that never receives a gradient (a branch that is not
            # exercised) must be named in `unused` like proj_attn; say so loudly instead of silently decaying it.
            self._checked_unused = True
            idle = [n for n, p in self.params if p._mig_slot[0] < self.used_numel and id(p) not in self._seen]
            if idle:
                import warnings
                warnings.warn("FlatAdamW: no gradient reached " + ", ".join(idle[:6]) + (" ..." if len(idle) > 6 else "") +
                              " in the first step; unlike torch.optim.AdamW (which skips grad=None) these parameters still "
                              "receive decoupled weight decay here. List their name fragments in `unused=` to exclude them.")
        if self.buckets is not None:
            self.finish_grad_sync()
        self.step_count += 1          # host mirror; the kernel reads the device counter (valid under graph replay)
        self.step_dev.add_(1)
        st = ops._stream()
        hp = (float(self.lr), float(self.betas[0]), float(self.betas[1]), float(self.eps), float(self.weight_decay),
              int(self.step_count))
        sumsq_ptr = None
        max_norm = 0.0
        if not (self.sharded and self.buckets is not None and self.buckets.nb):
            if self.max_grad_norm:
                call("mig_sumsq", ops._ptr(self.grad), ops._ptr(self.sumsq), ops._ptr(self._partials), self.used_numel, st)
                sumsq_ptr, max_norm = ops._ptr(self.sumsq), float(self.max_grad_norm)
            call("mig_adamw_step", ops._ptr(self.master), ops._ptr(self.grad), ops._ptr(self.m), ops._ptr(self.v),
                 self.used_numel, *hp, sumsq_ptr, max_norm, ops._ptr(self.shadow), ops._ptr(self.step_dev), st)
            return
        # ---- sharded: this rank owns slice `rank` of every bucket of the sharded region ----
        bk = self.buckets
        first = bk.rank * bk.piece
        own = lambda buf: ops._ptr(buf[first:])  # noqa: E731   (first owned element; pieces lie `bucket` apart)
        n_repl = self.used_numel - self.repl_lo
        if self.max_grad_norm:
            # ||g||^2 = sum over ranks of the owned slices (one scalar all-reduce; the same value on every rank, so the
            # clip coefficient and with it the replicas stay bit-identical) + the replicated region
            call("mig_sumsq_strided", own(self.grad), ops._ptr(self._ss[0:1]), ops._ptr(self._partials), bk.piece,
                 bk.bucket, bk.nb, st)
            dist.all_reduce(self._ss[0:1], op=dist.ReduceOp.SUM, group=bk.group)
            if n_repl:
                call("mig_sumsq", ops._ptr(self.grad[self.repl_lo:]), ops._ptr(self._ss[1:2]), ops._ptr(self._partials),
                     n_repl, st)
            else:
                self._ss[1:2].zero_()
            call("mig_add", 0, ops._ptr(self._ss[0:1]), ops._ptr(self._ss[1:2]), ops._ptr(self.sumsq), 1, st)
            sumsq_ptr, max_norm = ops._ptr(self.sumsq), float(self.max_grad_norm)
        call("mig_adamw_step_strided", own(self.master), own(self.grad), own(self.m), own(self.v), bk.piece, bk.bucket,
             bk.nb, *hp, sumsq_ptr, max_norm, own(self.shadow), ops._ptr(self.step_dev), st)
        if n_repl:
            lo = self.repl_lo
            call("mig_adamw_step", ops._ptr(self.master[lo:]), ops._ptr(self.grad[lo:]), ops._ptr(self.m[lo:]),
                 ops._ptr(self.v[lo:]), n_repl, *hp, sumsq_ptr, max_norm, ops._ptr(self.shadow[lo:]),
                 ops._ptr(self.step_dev), st)
        bk.gather(self.shadow)      # every rank needs all bf16 weights for the next forward: 2 bytes per parameter

    def consolidate(self) -> None:
        """COLLECTIVE. With the sharded update a rank holds current fp32 masters / Adam moments only for the slices it
        owns (the bf16 shadows are always complete). Call this on every rank before `module.state_dict()` /
        `torch_state_dict()` (checkpoints, train_ldm.py:466-491): it all-gathers master, m and v."""
        if self.sharded and self.buckets is not None and self.buckets.nb:
            for buf in (self.master, self.m, self.v):
                self.buckets.gather(buf)

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
            if off >= self.used_numel or self.step_count == 0 or id(p) not in self._seen:
                continue     # (torch keeps no state for parameters that never had a gradient)
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
                 graph_warmup_steps: int = 3, grad_accumulate_step: int = 1, shard_optimizer: Optional[bool] = None):
        self.unet, self.scheduler = unet, scheduler
        # config['grad_accumulate_step'] (train_ldm.py:173): the optimiser steps every k-th call of step(); gradients
        # of the micro-steps are SUMMED (the reference does not rescale the loss)
        self.grad_accumulate_step = max(1, int(grad_accumulate_step))
        self._micro = 0
        # data parallel + bf16: the optimiser is sharded by default (reduce-scatter, 1/world of AdamW per rank,
        # all-gather of the bf16 weights); shard_optimizer=False keeps the replicated all-reduce scheme
        self.opt = FlatAdamW(unet, lr=lr, weight_decay=weight_decay, max_grad_norm=grad_clip_max_norm,
                             bucket_mb=bucket_mb, shard=shard_optimizer)
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

    def checkpoint(self, epoch: int, validation_loss: float) -> dict:
        """COLLECTIVE (call on every rank; save on one). The reference's checkpoint dict (train_ldm.py:472-477):
        `network_state_dict` = the module's state_dict, `optimizer_state_dict` in torch.optim.AdamW's own layout. With
        the sharded optimiser the fp32 masters / moments are gathered from their owners first."""
        self.opt.consolidate()
        return {"epoch": epoch, "network_state_dict": self.unet.state_dict(),
                "optimizer_state_dict": self.opt.torch_state_dict(), "validation_loss": validation_loss}


@torch.no_grad()
def latent_scale_factor(autoencoder, images: torch.Tensor, src: int = 0) -> torch.Tensor:
    """`scale_factor = 1 / std(z)` of the FIRST batch (train_ldm.py:98-118), computed ONCE on rank `src` and broadcast:
    it is data dependent, and every data-parallel rank must scale its latents by the same number.
    Returns a 0-d tensor on the autoencoder's device."""
    dev = next(autoencoder.parameters()).device
    sf = torch.zeros((), dtype=torch.float32, device=dev)
    rank = dist.get_rank() if _dist_on() else src
    if rank == src:
        z = autoencoder.encode_stage_2_inputs(images.to(dev))
        sf = (1.0 / torch.std(z.float())).reshape(())
    if _dist_on():
        dist.broadcast(sf, src=src)
    return sf


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
