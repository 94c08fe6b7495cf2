"""Kernel-level breakdown of one LDM training step (torch.profiler / CUPTI): which kernels own the step time."""
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, ".")
import bench  # noqa: E402
import medical_image_generation_b200 as mig  # noqa: E402
from medical_image_generation_b200 import planner  # noqa: E402
from medical_image_generation_b200.engine import LDMTrainer  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
torch.manual_seed(0)
model = bench.rerandomize_zero_init(mig.DiffusionModelUNet(**bench.unet_kwargs())).cuda().train()
tr = LDMTrainer(model, mig.DDPMScheduler(**planner.LDM_SCHEDULER_KWARGS))
x = torch.randn(B, *bench.LATENT, device="cuda")
for _ in range(3):
    tr.step(x)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(2):
        tr.step(x)
    torch.cuda.synchronize()
ev = prof.key_averages()
def _is_kernel(k):
    return k.startswith(("void ", "mig::", "Memset", "Memcpy", "nccl"))


rows = sorted(((e.device_time_total, e.count, e.key) for e in ev if e.device_time_total > 0 and _is_kernel(e.key)),
              reverse=True)
tot = sum(r[0] for r in rows)
print(f"total device time per step: {tot / 2 / 1e3:.2f} ms")
for t, c, k in rows[:45]:
    print(f"{t / 2 / 1e3:9.3f} ms {100 * t / tot:5.1f}%  x{c // 2:<5d} {k[:110]}")
cpu = sorted(((e.self_cpu_time_total, e.count, e.key) for e in ev), reverse=True)[:12]
print("-- host side (self CPU time per step)")
for t, c, k in cpu:
    print(f"{t / 2 / 1e3:9.3f} ms x{c // 2:<5d} {k[:100]}")
