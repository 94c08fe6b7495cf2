"""Kernel-level breakdown of one DDPM reverse step of BASELINE config 4 (candidate A), torch.profiler / CUPTI."""
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, ".")
import medical_image_generation_b200 as mig  # noqa: E402
from medical_image_generation_b200 import planner  # noqa: E402

widths = (32, 64, 128) if (len(sys.argv) < 2 or sys.argv[1] == "A") else (64, 128, 256)
torch.manual_seed(0)
cfg = dict(spatial_dims=3, in_channels=1, out_channels=1, num_res_blocks=2, num_channels=list(widths),
           attention_levels=[False, False, True], num_head_channels=[0, 0, widths[2]], norm_num_groups=32,
           strides=[[1, 1, 1], [2, 2, 1], [2, 2, 2]], kernel_sizes=[[3, 3, 3]] * 3, paddings=[[1, 1, 1]] * 3)
m = mig.DiffusionModelUNet(**cfg).cuda().eval()
with torch.no_grad():
    for p in m.parameters():
        if float(p.abs().max()) == 0:
            p.normal_(0, 0.02)
s = mig.DDPMScheduler(**planner.LDM_SCHEDULER_KWARGS)
s.noise_mode = "device"
s.set_timesteps(1000)
img = torch.randn(1, 1, 128, 128, 64, device="cuda")
ts = s.timesteps[:6]
with torch.no_grad():
    for t in ts[:3]:
        img, _ = s.step(m(img, timesteps=torch.Tensor((t,)).cuda()), t, img)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for t in ts[3:]:
            img, _ = s.step(m(img, timesteps=torch.Tensor((t,)).cuda()), t, img)
        torch.cuda.synchronize()
n = len(ts) - 3
rows = sorted(((e.device_time_total, e.count, e.key) for e in prof.key_averages()
               if e.device_time_total > 0 and e.key.startswith(("void ", "mig::", "Memset", "Memcpy"))), reverse=True)
tot = sum(r[0] for r in rows)
print(f"total device time per reverse step: {tot / n / 1e3:.2f} ms")
for t, c, k in rows[:30]:
    print(f"{t / n / 1e3:9.3f} ms {100 * t / tot:5.1f}%  x{c // n:<5d} {k[:120]}")
