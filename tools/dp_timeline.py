"""Timeline of one data-parallel training step (torch.profiler on rank 0): when each NCCL all-reduce runs relative to
the compute stream. Launch: python -m torch.distributed.run --nproc-per-node 2 tools/dp_timeline.py"""
import os
import sys

import torch
import torch.distributed as dist
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, ".")
import bench  # noqa: E402
import medical_image_generation_b200 as mig  # noqa: E402
from medical_image_generation_b200 import planner  # noqa: E402
from medical_image_generation_b200.engine import LDMTrainer  # noqa: E402

rank = int(os.environ.get("RANK", "0"))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
dist.init_process_group("nccl")
torch.manual_seed(0)
model = bench.rerandomize_zero_init(mig.DiffusionModelUNet(**bench.unet_kwargs())).cuda().train()
tr = LDMTrainer(model, mig.DDPMScheduler(**planner.LDM_SCHEDULER_KWARGS))
x = torch.randn(8, *bench.LATENT, device="cuda")
for _ in range(3):
    tr.step(x)
torch.cuda.synchronize()
dist.barrier()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    tr.step(x)
    torch.cuda.synchronize()
if rank == 0:
    ev = [e for e in prof.events() if e.device_type is not None and str(e.device_type).endswith("CUDA") and e.device_time > 0]
    ev.sort(key=lambda e: e.time_range.start)
    t0 = ev[0].time_range.start
    end = max(e.time_range.end for e in ev)
    print(f"step span {(end - t0) / 1e3:.2f} ms, {len(ev)} device events")
    nccl = [e for e in ev if "nccl" in e.name.lower()]
    comp = [e for e in ev if "nccl" not in e.name.lower()]
    print(f"NCCL kernels: {len(nccl)}, total {sum(e.device_time for e in nccl) / 1e3:.2f} ms; "
          f"first starts at {(nccl[0].time_range.start - t0) / 1e3:.2f} ms, last ends at {(nccl[-1].time_range.end - t0) / 1e3:.2f} ms")
    last_wgrad = max((e for e in comp if "wgrad" in e.name), key=lambda e: e.time_range.end)
    adam = [e for e in comp if "adamw" in e.name][0]
    sumsq = [e for e in comp if "sumsq" in e.name][0]
    print(f"last wgrad kernel ends at {(last_wgrad.time_range.end - t0) / 1e3:.2f} ms; sumsq starts {(sumsq.time_range.start - t0) / 1e3:.2f} ms; "
          f"adamw {(adam.time_range.start - t0) / 1e3:.2f} -> {(adam.time_range.end - t0) / 1e3:.2f} ms")
    busy = sum(e.device_time for e in comp) / 1e3
    print(f"compute-stream kernel time {busy:.2f} ms")
    import collections
    agg = collections.defaultdict(lambda: [0, 0.0])
    for e in comp:
        k = e.name.split("(")[0][:70]
        agg[k][0] += 1
        agg[k][1] += e.device_time
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
        print(f"   {t / 1e3:8.3f} ms x{c:<4d} {k}")
    print(f"adamw kernels: {[(round((e.time_range.start - t0) / 1e3, 2), round(e.device_time / 1e3, 3)) for e in comp if 'adamw' in e.name]}")
    for e in nccl:
        print(f"   {e.name[:48]:48s} start {(e.time_range.start - t0) / 1e3:7.2f} ms dur {e.device_time / 1e3:6.3f} ms")
dist.barrier()
torch.cuda.synchronize()
os._exit(0)
