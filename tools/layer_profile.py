"""Per-layer-shape time of every conv / linear implicit-GEMM call inside real LDM training steps (CUDA events)."""
import sys

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
import medical_image_generation_b200 as mig  # noqa: E402
from medical_image_generation_b200 import ops, planner  # noqa: E402
from medical_image_generation_b200.engine import LDMTrainer  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
torch.manual_seed(0)
model = bench.rerandomize_zero_init(mig.DiffusionModelUNet(**bench.unet_kwargs())).cuda().train()
tr = LDMTrainer(model, mig.DDPMScheduler(**planner.LDM_SCHEDULER_KWARGS))
x = torch.randn(B, *bench.LATENT, device="cuda")
for _ in range(3):
    tr.step(x)
torch.cuda.synchronize()
steps = 3
ops.profile_start()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    tr.step(x)
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
print(f"step {total_ms:.2f} ms; conv/linear calls {conv_ms:.2f} ms ({100*conv_ms/total_ms:.0f}%)")
for (kind, shape), (f, ms, n) in rows[:40]:
    print(f"{ms:7.3f} ms  x{n // steps:<3d} {kind:5s} Cin={shape[0]:<5d} Cout={shape[1]:<5d} out={shape[2]} k={shape[3]}  "
          f"{f / ms / 1e9 if ms > 0 else 0:7.0f} TF/s")
bykind = {}
for (kind, shape), (f, ms, n) in agg.items():
    d = bykind.setdefault(kind, [0.0, 0.0]); d[0] += f; d[1] += ms
print({k: (round(v[1], 2), round(v[0] / v[1] / 1e9)) for k, v in bykind.items()})
