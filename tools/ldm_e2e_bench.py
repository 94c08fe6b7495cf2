"""BASELINE config 5: 3D LDM end-to-end training step on one GPU's share: frozen AutoencoderKL (32,64,128), in/out 2
channels, latent 3) encodes B x 2x160x160x128 volumes under no_grad (train_ldm.py:134,155), the latents (3x40x40x32,
scaled) go through the LDM-default U-Net training step (noise, add_noise, fwd, MSE, bwd, clip, fused AdamW).
Synthetic data, random-init weights, bf16. Prints samples/s and the per-layer conv profile of one step.
usage: python tools/ldm_e2e_bench.py [B=2] [steps=5]"""
import sys

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
import medical_image_generation_b200 as mig  # noqa: E402
from medical_image_generation_b200 import ops, planner  # noqa: E402
from medical_image_generation_b200.engine import LDMTrainer  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
vol = (160, 160, 128)
torch.manual_seed(0)
ae_kw = planner.autoencoder_kwargs(vol, in_channels=2, latent_channels=3)
ae = mig.AutoencoderKL(**ae_kw).cuda().eval().requires_grad_(False)
lat = planner.compute_output_size(vol, ae_kw["downsample_parameters"])
unet = bench.rerandomize_zero_init(mig.DiffusionModelUNet(**planner.ddpm_kwargs(lat, latent_channels=3))).cuda().train()
tr = LDMTrainer(unet, mig.DDPMScheduler(**planner.LDM_SCHEDULER_KWARGS), lr=2e-5, grad_clip_max_norm=1.0)
x = torch.rand(B, 2, *vol, device="cuda")
with torch.no_grad():
    z0 = ae.encode_stage_2_inputs(x)
    scale = 1.0 / float(z0.float().std())   # train_ldm.py:110-112
print(f"volumes {tuple(x.shape)} -> latents {tuple(z0.shape)}; scale_factor {scale:.4f}; "
      f"U-Net params {sum(p.numel() for p in unet.parameters()) / 1e6:.1f} M", flush=True)


def step():
    with torch.no_grad():
        z = ae.encode_stage_2_inputs(x) * scale
    return tr.step(z)


for _ in range(2):
    loss = step()
torch.cuda.synchronize()
e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
e[0].record()
with torch.no_grad():
    for _ in range(steps):
        z = ae.encode_stage_2_inputs(x) * scale
e[1].record()
ops.profile_start()
for _ in range(steps):
    loss = step()
e[2].record()
torch.cuda.synchronize()
prof = ops.profile_stop()
enc_ms = e[0].elapsed_time(e[1]) / steps
ms = e[1].elapsed_time(e[2]) / steps
flop = 18.13e12 * B
print(f"config 5, B={B}: {ms:.1f} ms/step (AE encode alone {enc_ms:.1f} ms) -> {B / ms * 1e3:.2f} samples/s; "
      f"{flop / ms / 1e9:.0f} TFLOP/s of algorithmic 18.13 TFLOP/sample; loss {float(loss.detach()):.4f}; "
      f"peak memory {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
agg = {}
for kind, flops, shape, a, b in prof:
    d = agg.setdefault((kind, shape), [0.0, 0.0, 0])
    d[0] += flops; d[1] += a.elapsed_time(b); d[2] += 1
tot = sum(v[1] for v in agg.values()) / steps
print(f"conv / linear calls {tot:.1f} ms ({100 * tot / ms:.0f}%)")
for (kind, shape), v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:24]:
    print(f"  {v[1] / steps:7.3f} ms x{v[2] // steps:<3d} {kind:5s} {shape}  {v[0] / v[1] / 1e9:6.0f} TF/s")
