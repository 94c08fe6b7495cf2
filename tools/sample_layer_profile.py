"""Per-layer-shape time of every conv / linear call inside DDPM reverse steps of BASELINE config 4 (candidate A)."""
import sys

import torch

sys.path.insert(0, ".")
import medical_image_generation_b200 as mig  # noqa: E402
from medical_image_generation_b200 import ops, planner  # noqa: E402

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
steps = 3
with torch.no_grad():
    for t in ts[:3]:
        img, _ = s.step(m(img, timesteps=torch.Tensor((t,)).cuda()), t, img)
    torch.cuda.synchronize()
    ops.profile_start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for t in ts[3:]:
        img, _ = s.step(m(img, timesteps=torch.Tensor((t,)).cuda()), t, img)
    e1.record()
    torch.cuda.synchronize()
    prof = ops.profile_stop()
total_ms = e0.elapsed_time(e1) / steps
agg = {}
for kind, flops, shape, a, b in prof:
    d = agg.setdefault((kind, shape), [0.0, 0.0, 0])
    d[0] += flops / steps
    d[1] += a.elapsed_time(b) / steps
    d[2] += 1
rows = sorted(agg.items(), key=lambda kv: -kv[1][1])
conv_ms = sum(v[1] for v in agg.values())
print(f"reverse step {total_ms:.2f} ms (eager); conv/linear calls {conv_ms:.2f} ms ({100*conv_ms/total_ms:.0f}%)")
for (kind, shape), (f, ms, n) in rows[:40]:
    print(f"{ms:7.3f} ms  x{n // steps:<3d} {kind:5s} Cin={shape[0]:<5d} Cout={shape[1]:<5d} out={shape[2]} k={shape[3]}  "
          f"{f / ms / 1e9 if ms > 0 else 0:7.0f} TF/s")
