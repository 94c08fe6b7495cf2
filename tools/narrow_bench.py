"""Forward-only micro-benchmark of the narrow-channel layers of the config-4 sampling U-Net (candidate A) -- CUDA events.
usage: python tools/narrow_bench.py [reps] [only-index]"""
import sys

import torch

sys.path.insert(0, ".")
from medical_image_generation_b200 import ops  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
only = int(sys.argv[2]) if len(sys.argv) > 2 else -1
SHAPES = [  # (Cin, Cout, (D, H, W), k)
    (64, 32, (128, 128, 64), 1), (96, 32, (128, 128, 64), 1), (64, 32, (128, 128, 64), 3), (32, 32, (128, 128, 64), 3),
    (32, 1, (128, 128, 64), 3), (1, 32, (128, 128, 64), 3), (64, 64, (64, 64, 64), 3), (128, 128, (32, 32, 32), 3),
    (128, 128, (1, 1, 32768), 1), (192, 64, (64, 64, 64), 1)]
print(torch.cuda.get_device_name(0))
for i, (Cin, Cout, sp, k) in enumerate(SHAPES):
    if only >= 0 and i != only:
        continue
    x = torch.randn(1, Cin, *sp, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last_3d)
    w = (torch.randn(Cout, Cin, k, k, k, device="cuda") * 0.05).contiguous(memory_format=torch.channels_last_3d)
    b = torch.zeros(Cout, device="cuda")
    vox = sp[0] * sp[1] * sp[2]
    flops = 2.0 * vox * Cout * Cin * k ** 3
    bytes_ = 2.0 * vox * (Cin + Cout)
    with torch.no_grad():
        for _ in range(2):
            y = ops.conv_nd(x, w, b, 1, k // 2)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            y = ops.conv_nd(x, w, b, 1, k // 2)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"[{i}] {Cin:4d}->{Cout:<4d} {sp} k={k}: {ms * 1e3:8.1f} us  {flops / ms / 1e9:7.0f} TF/s  {bytes_ / ms / 1e6:7.0f} GB/s algorithmic",
          flush=True)

# GroupNorm(32 groups) + SiLU forward on the same tensors (algorithmic bytes: read x + write y)
if only < 0:
    for C, sp in ((32, (128, 128, 64)), (64, (128, 128, 64)), (96, (128, 128, 64)), (64, (64, 64, 64)), (128, (64, 64, 64)),
                  (128, (32, 32, 32)), (256, (32, 32, 32))):
        x = torch.randn(1, C, *sp, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last_3d)
        g, b = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
        with torch.no_grad():
            for _ in range(2):
                y = ops.group_norm(x, g, b, 32, 1e-6, silu=True)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                y = ops.group_norm(x, g, b, 32, 1e-6, silu=True)
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        nbytes = 4.0 * x.numel()
        print(f"GN+SiLU C={C:<4d} {sp}: {ms * 1e3:8.1f} us  {nbytes / ms / 1e6:7.0f} GB/s algorithmic (x read twice: {1.5 * nbytes / ms / 1e6:.0f} GB/s moved)",
              flush=True)
