#!/usr/bin/env python
"""Headline benchmark: 3D LDM U-Net training samples/s (BASELINE.json config 3).

    python bench.py --gpus N --steps K --warmup W            # B200 arm (ours)
    python bench.py --impl reference --gpus N --steps K ...  # reference's CPU path on the host cores (oracle port)

Workload ("ldm_unet_train_3x24x24x24"): LDM-default DiffusionModelUNet (create_ddpm_dict, configuration.py:865-902:
widths 256/512/768, attention on the two coarse levels with one head of 512/768 channels, 441 M parameters) on
3x24x24x24 latents, batch 8 PER GPU (weak scaling), epsilon-prediction MSE, global-norm clip 1.0, AdamW lr 2e-5 --
one optimiser step of train_ldm.py:143-183 per bench step. Synthetic latents, random-init weights (the zero-
initialised convs are re-randomised so no layer is a no-op).

One JSON line is printed by rank 0; see the repo task description for the field contract.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LATENT = (3, 24, 24, 24)
METRIC = "3D LDM U-Net train samples/s"
WORKLOAD = "ldm_unet_train_3x24x24x24 (BASELINE.json configs[2]: LDM-default U-Net 256/512/768, batch 8 per GPU)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=8, help="per-GPU batch")
    ap.add_argument("--cpu-steps", type=int, default=2, help="CPU baseline steps at batch 1 (bounded sample)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--engine", type=int, default=0, help="0 auto (tcgen05 where eligible), 1 SIMT only")
    ap.add_argument("--no-graph", action="store_true", help="launch the step eagerly instead of replaying a CUDA graph")
    ap.add_argument("--bucket-mb", type=float, default=64.0, help="gradient all-reduce bucket size (multi-GPU)")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(tflops=p.get("bf16_tflops_sustained", p.get("bf16_tflops")), burst=p.get("bf16_tflops"),
                    hbm=p.get("hbm_gbs"), source="measured (MEASURED_PEAKS.json, sustained bf16)")
    return dict(tflops=1400.0, burst=1590.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


# --------------------------------------------------------------------------------------------------
# model construction shared by both arms
# --------------------------------------------------------------------------------------------------
def unet_kwargs():
    from medical_image_generation_b200 import planner
    return planner.ddpm_kwargs(list(LATENT[1:]), latent_channels=LATENT[0])


def rerandomize_zero_init(module, seed=1234, std=0.02):
    import torch
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for _, p in sorted(module.named_parameters()):
            if p.numel() and float(p.abs().max()) == 0.0:
                p.copy_(torch.randn(p.shape, generator=g) * std)
    return module


# --------------------------------------------------------------------------------------------------
# CPU baseline: the oracle port (the reference's algorithm restated in plain torch CPU fp32)
# --------------------------------------------------------------------------------------------------
def cpu_reference_steps(steps: int, warmup: int, seed: int = 0):
    """Times `steps` optimiser steps at batch 1 on the host cores. Returns (samples/s, cores, seconds per step)."""
    import torch
    from oracle import torch_oracle as O
    from oracle.ddpm_oracle import OracleDDPMScheduler
    import medical_image_generation_b200 as mig
    from medical_image_generation_b200 import planner
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(seed)
    cfg = unet_kwargs()
    shapes_model = rerandomize_zero_init(mig.DiffusionModelUNet(**cfg))
    params = {k: v.detach().clone().contiguous().requires_grad_(True) for k, v in shapes_model.state_dict().items()}
    del shapes_model
    used = [v for k, v in params.items() if "proj_attn" not in k]
    opt = torch.optim.AdamW(used, lr=2e-5)
    sched = OracleDDPMScheduler(**planner.LDM_SCHEDULER_KWARGS)
    times = []
    for i in range(warmup + steps):
        x0 = torch.randn(1, *LATENT)
        noise = torch.randn_like(x0)
        t = torch.randint(0, 1000, (1,))
        t0 = time.perf_counter()
        pred = O.unet_forward(params, cfg, sched.add_noise(x0, noise, t), t)
        loss = torch.nn.functional.mse_loss(pred.float(), noise.float())
        opt.zero_grad(set_to_none=True)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(used, 1.0)
        opt.step()
        float(loss)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    sec = sum(times) / len(times)
    return 1.0 / sec, cores, sec


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sps, cores, sec = cpu_reference_steps(args.steps, args.warmup)
    sample = (f"{args.steps} timed optimiser steps at batch 1 (1/{args.batch} of one GPU's batch) of the same U-Net, "
              f"fp32, {cores} host threads, oracle port of the reference modules")
    line = {"impl": "reference", "metric": METRIC, "value": sps, "unit": "samples/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "per_step_sample": "batch 1 on CPU"},
            "cpu_baseline": {"value": sps, "unit": "samples/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": sps, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# clocks sampler
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, row in self.rows:
            if ts < t0 or ts > t1 + 0.2:
                continue
            parts = [p.strip() for p in row.split(",")]
            try:
                sm.append(float(parts[0]))
                mx = float(parts[1])
                for n, v in zip(names, parts[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:  # noqa: BLE001
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# --------------------------------------------------------------------------------------------------
# B200 arm
# --------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    import medical_image_generation_b200 as mig
    from medical_image_generation_b200 import _lib, ops, planner
    from medical_image_generation_b200.engine import LDMTrainer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (b200 arm) needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    ops.set_engine(args.engine)

    torch.manual_seed(0)
    cfg = unet_kwargs()
    model = rerandomize_zero_init(mig.DiffusionModelUNet(**cfg, compute_dtype=torch.bfloat16)).to(dev).train()
    n_params = sum(p.numel() for p in model.parameters())
    sched = mig.DDPMScheduler(**planner.LDM_SCHEDULER_KWARGS)
    # whole-step CUDA graph: two eager steps, capture on the third untimed step, replay afterwards
    trainer = LDMTrainer(model, sched, lr=2e-5, grad_clip_max_norm=1.0, cuda_graph=not args.no_graph, bucket_mb=args.bucket_mb,
                         graph_warmup_steps=2)
    B = args.batch
    gen = torch.Generator().manual_seed(1000 + rank)
    host = [torch.randn((B, *LATENT), generator=gen).pin_memory() for _ in range(4)]   # rank-offset seeds
    resident = host[0].to(dev)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- warm-up (also compiles nothing: kernels are prebuilt; this warms allocator + L2 + clocks) ----
    for i in range(max(args.warmup, 4)):   # >= 4 untimed steps so that graph capture never lands in the timed region
        trainer.step(resident)
    sync_all()

    # ---- timed region 1: device-resident inputs -> `value` ----
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    sync_all()
    w0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        loss = trainer.step(resident)
    e1.record()
    sync_all()
    w1 = time.perf_counter()
    ms = e0.elapsed_time(e1)
    # per-kernel CUDA events need the Python launch path (a graph replay has no per-node events): the same step is
    # run eagerly, instrumented, directly after the timed region -- same workload, same kernels, same stream.
    prof_steps = min(args.steps, 3)
    for i in range(2):   # graph capture empties the caching allocator: re-warm the eager path first
        trainer._eager_step(resident)
    sync_all()
    ops.profile_start()
    launches0 = _lib.launch_count
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    for i in range(prof_steps):
        trainer._eager_step(resident)
    p1.record()
    sync_all()
    launches = (_lib.launch_count - launches0) // prof_steps * args.steps
    prof = ops.profile_stop()
    eager_ms_per_step = p0.elapsed_time(p1) / prof_steps
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t)
    clocks = sampler.stop(w0, w1) if rank == 0 else None
    final_loss = float(loss.detach())

    # ---- timed region 2: end to end through the public API with HOST (pinned) inputs -> `e2e` ----
    sync_all()
    t0 = time.perf_counter()
    for i in range(args.steps):
        x = host[i % len(host)].to(dev, non_blocking=True)          # H2D of this step's latents
        loss = trainer.step(x)
        _ = loss.item()                                               # D2H of the step's loss (train_ldm.py:182)
    sync_all()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te)

    if rank == 0:
        pk = peaks()
        # roofline of the dominant kernel family: tcgen05 implicit-GEMM conv launches, timed live with CUDA events
        agg = {}
        for kind, flops, shape, a, b in prof:
            d = agg.setdefault(kind, [0.0, 0.0, 0])
            d[0] += flops
            d[1] += a.elapsed_time(b) * 1e-3
            d[2] += 1
        detail = {k: {"tflops": v[0] / v[1] / 1e12 if v[1] > 0 else None, "launches": v[2], "seconds": v[1],
                      "flop": v[0]} for k, v in agg.items()}
        gemm_f = sum(v[0] for k, v in agg.items() if k in ("fwd", "dgrad"))
        gemm_s = sum(v[1] for k, v in agg.items() if k in ("fwd", "dgrad"))
        gemm_n = sum(v[2] for k, v in agg.items() if k in ("fwd", "dgrad"))
        all_f, all_s = sum(v[0] for v in agg.values()), sum(v[1] for v in agg.values())
        achieved = gemm_f / gemm_s / 1e12 if gemm_s > 0 else 0.0
        # the single largest launch family (256->256, 3x3x3, 8 x 24^3: 14 fwd+dgrad launches per step), timed live; its
        # DRAM traffic comes from the committed ncu --set full capture of exactly this launch
        big = [(f, a.elapsed_time(b)) for kind, f, shape, a, b in prof
               if kind in ("fwd", "dgrad") and tuple(shape[:2]) == (256, 256) and tuple(shape[2]) == (24, 24, 24)
               and tuple(shape[3]) == (3, 3, 3)]
        sampled = None
        if big:
            sampled = {"layer": "Conv3d 256->256 3x3x3 on 8x24^3 voxels (fwd / dgrad)", "launches": len(big),
                       "flop_per_launch": big[0][0], "avg_launch_ms": sum(t for _, t in big) / len(big),
                       "tflops": sum(f for f, _ in big) / sum(t for _, t in big) / 1e9,
                       "algorithmic_bytes_per_launch": 2 * B * 24 ** 3 * 256 * 2 + 256 * 27 * 256 * 2,
                       "ncu_dram_bytes_per_launch": 73.9e6,
                       "ncu_source": "profiles/r01_ncu_full_conv256_24cube.csv (60.3 MB read + 13.6 MB written: the "
                                     "output tile stays in the 126 MB L2 for the next kernel)"}
        roofline = {"bound": "tensor", "kernel": "conv_tma_kernel (persistent tcgen05 implicit-GEMM conv fwd+dgrad)",
                    "achieved": achieved, "peak": pk["tflops"], "unit": "TFLOP/s", "frac": achieved / pk["tflops"],
                    "traffic": sampled["ncu_dram_bytes_per_launch"] if sampled else None,
                    "traffic_of": "the sampled launch below (per-launch traffic differs by layer shape)",
                    "sampled_launch": sampled, "peak_source": pk["source"],
                    "flop_per_launch": gemm_f / max(gemm_n, 1), "avg_launch_ms": gemm_s / max(gemm_n, 1) * 1e3,
                    "detail": detail,
                    "timed_over": f"{prof_steps} eager instrumented steps directly after the timed region",
                    "all_conv": {"tflops": all_f / all_s / 1e12 if all_s else None,
                                 "share_of_step": all_s / (eager_ms_per_step * prof_steps * 1e-3)},
                    "eager_ms_per_step": eager_ms_per_step,
                    "whole_step_model_tflops": 4.302e12 * B * args.steps / (ms_total * 1e-3) / 1e12}
        cpu = None
        if not args.no_cpu_baseline and world == 1:   # the contract: rank 0 at N=1 only
            sps, cores, sec = cpu_reference_steps(args.cpu_steps, 0)
            cpu = {"value": sps, "unit": "samples/s", "cores": cores, "kind": "port",
                   "sample": f"{args.cpu_steps} optimiser steps at batch 1 of the same U-Net (oracle port, fp32, "
                             f"{cores} threads): {sec:.2f} s/step"}
        value = world * B * args.steps / (ms_total * 1e-3)
        line = {"metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": {"workload": WORKLOAD, "per_gpu_batch": B, "global_batch": B * world,
                           "latent": list(LATENT), "params": n_params, "parallelism": f"dp{world}",
                           "optimizer": "AdamW lr 2e-5, clip 1.0 (fused flat)", "final_loss": final_loss,
                           "l2": "no explicit flush: per-step working set (0.88 GB bf16 filters + >10 GB "
                                 "activations) is far larger than the 126 MB L2"},
                "e2e": {"value": world * B * args.steps / e2e_s, "unit": "samples/s",
                        "h2d_bytes_per_step": B * 4 * LATENT[0] * LATENT[1] * LATENT[2] * LATENT[3],
                        "d2h_bytes_per_step": 4},
                "gpu_launches": launches,
                "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu}
        print(json.dumps(line), flush=True)
    if world > 1:
        # Teardown: the captured step graph holds NCCL work; destroying the communicator with such a graph alive can
        # block for minutes. Drop the graph, drain the device, meet at a barrier and leave without the interpreter's
        # teardown (every result is already printed and flushed).
        trainer._graph = None
        import gc
        gc.collect()
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def main():
    args = parse()
    # failsafe: a wedged collective must not hold the GPUs until the caller's limit
    limit = float(os.environ.get("MIG_BENCH_LIMIT_S", "1500"))
    watchdog = threading.Timer(limit, lambda: (sys.stderr.write("bench.py: time limit reached, aborting\n"), os._exit(3)))
    watchdog.daemon = True
    watchdog.start()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
