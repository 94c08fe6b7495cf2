"""Kernel-level breakdown of one AutoencoderKL training step (BASELINE config 2), torch.profiler / CUPTI."""
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, ".")
import medical_image_generation_b200 as mig  # noqa: E402
from medical_image_generation_b200 import planner  # noqa: E402
from medical_image_generation_b200.engine import AETrainer  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
size = int(sys.argv[2]) if len(sys.argv) > 2 else 96
torch.manual_seed(0)
ae = mig.AutoencoderKL(**planner.autoencoder_kwargs([size] * 3, in_channels=1, latent_channels=3, levels=2)).cuda().train()
tr = AETrainer(ae, lr=5e-5)
x = torch.rand(B, 1, size, size, size, device="cuda")
for _ in range(3):
    tr.step(x)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(2):
        tr.step(x)
    torch.cuda.synchronize()
rows = sorted(((e.device_time_total, e.count, e.key) for e in prof.key_averages()
               if e.device_time_total > 0 and e.key.startswith(("void ", "mig::", "Memset", "Memcpy"))), reverse=True)
tot = sum(r[0] for r in rows)
print(f"total device time per step: {tot / 2 / 1e3:.2f} ms")
for t, c, k in rows[:28]:
    print(f"{t / 2 / 1e3:9.3f} ms {100 * t / tot:5.1f}%  x{c // 2:<5d} {k[:120]}")
