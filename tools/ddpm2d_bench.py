"""BASELINE config 1: 2-D DDPM training on 64x64 images (the reference's CPU-runnable smoke configuration, SURVEY
section 8d row 1): strided U-Net with the LDM-default widths (171 M parameters), batch 2, 10 AdamW(2e-5) steps with
clip 1.0 on torch.rand(2,1,64,64) (seed 42). Reference on 8 CPU cores: 33.5 s for the 10 steps, loss 0.99 -> 0.92."""
import sys
import time

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
import medical_image_generation_b200 as mig  # noqa: E402
from medical_image_generation_b200 import planner  # noqa: E402
from medical_image_generation_b200.engine import LDMTrainer  # noqa: E402

torch.manual_seed(42)
kw = planner.ddpm_kwargs([64, 64], latent_channels=1)
unet = bench.rerandomize_zero_init(mig.DiffusionModelUNet(**kw)).cuda().train()
print(f"2-D U-Net: {sum(p.numel() for p in unet.parameters()) / 1e6:.1f} M parameters")
tr = LDMTrainer(unet, mig.DDPMScheduler(**planner.LDM_SCHEDULER_KWARGS), lr=2e-5, grad_clip_max_norm=1.0, cuda_graph=True)
x = torch.rand(2, 1, 64, 64, device="cuda")
losses = []
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(10):
    losses.append(float(tr.step(x).detach()))
torch.cuda.synchronize()
t1 = time.perf_counter()
for _ in range(20):
    tr.step(x)
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"config 1: first 10 steps {t1 - t0:.2f} s (includes lazy initialisation); steady state {(t2 - t1) / 20 * 1e3:.1f} ms/step; "
      f"loss {losses[0]:.3f} -> {losses[-1]:.3f}")
