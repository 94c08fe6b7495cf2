"""BASELINE config 2: 3D AutoencoderKL (32,64,128), latent 3, train step (L1 + 1e-7*KL, Adam-style flat AdamW) on
synthetic 1x96^3 volumes, bf16, 1 GPU. Prints samples/s and the per-layer conv profile."""
import sys
import time

import torch

sys.path.insert(0, ".")
import medical_image_generation_b200 as mig  # noqa: E402
from medical_image_generation_b200 import ops, planner  # noqa: E402
from medical_image_generation_b200.engine import FlatAdamW  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
size = int(sys.argv[2]) if len(sys.argv) > 2 else 96
torch.manual_seed(0)
kw = planner.autoencoder_kwargs([size] * 3, in_channels=1, latent_channels=3, levels=2)
ae = mig.AutoencoderKL(**kw).cuda().train()
opt = FlatAdamW(ae, lr=5e-5, weight_decay=0.0, max_grad_norm=1.0, unused=())
x = torch.rand(B, 1, size, size, size, device="cuda")


def step():
    opt.zero_grad()
    recon, z_mu, z_sigma = ae(x)
    loss = ops.l1_loss(recon, x) + 1e-7 * ops.kl_loss(z_mu, z_sigma)
    loss.backward()
    opt.step()
    return loss


for _ in range(3):
    l = step()
torch.cuda.synchronize()
n = 5
ops.profile_start()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(n):
    l = step()
e1.record()
torch.cuda.synchronize()
prof = ops.profile_stop()
ms = e0.elapsed_time(e1) / n
print(f"AE train {size}^3 B={B}: {ms:.1f} ms/step -> {B / ms * 1e3:.2f} samples/s; loss {float(l):.4f}; "
      f"model {3.157e12 * B / ms / 1e9:.0f} TFLOP/s of algorithmic 3.157 TFLOP/sample")
agg = {}
for kind, f, shape, a, b in prof:
    d = agg.setdefault((kind, shape), [0.0, 0.0, 0]); d[0] += f / n; d[1] += a.elapsed_time(b) / n; d[2] += 1
conv_ms = sum(v[1] for v in agg.values())
print(f"conv calls {conv_ms:.1f} ms ({100 * conv_ms / ms:.0f}%)")
for (kind, shape), (f, t, c) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:16]:
    print(f"{t:7.3f} ms x{c // n:<2d} {kind:5s} Cin={shape[0]:<4d} Cout={shape[1]:<4d} out={shape[2]} k={shape[3]} {f / t / 1e9:6.0f} TF/s")
