"""Micro-benchmark of the conv implicit-GEMM kernels at the BASELINE config-3 layer shapes (CUDA-event timing).
usage: python tools/conv_bench.py [reps] [which=all|big]"""
import sys
import time

import torch

sys.path.insert(0, ".")
from medical_image_generation_b200 import ops  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
which = sys.argv[2] if len(sys.argv) > 2 else "all"
epi = len(sys.argv) > 3 and sys.argv[3] == "epi"     # forward with the fused epilogue inputs of a ResnetBlock's conv2
SHAPES = [  # (N, Cin, Cout, spatial) 3x3x3 stride 1 pad 1 -- the LDM-default U-Net at 3x24^3, batch 8
    (8, 256, 256, 24), (8, 512, 512, 24), (8, 768, 256, 24), (8, 512, 512, 12), (8, 1280, 512, 12),
    (8, 768, 768, 12), (8, 768, 768, 6), (8, 1536, 768, 6)]
if which == "big":
    SHAPES = SHAPES[:2]
elif which == "deep":
    SHAPES = SHAPES[3:]
elif which == "six":   # the 6^3 level only (split-K plans)
    SHAPES = [(8, 768, 768, 6)]
elif which == "mid":   # 64/128-channel layers of the AE / pixel-space U-Nets
    SHAPES = [(1, 64, 64, 64), (2, 64, 64, 48), (1, 128, 128, 32), (1, 64, 32, 96)]
print(torch.cuda.get_device_name(0))
for N, Cin, Cout, s in SHAPES:
    x = torch.randn(N, Cin, s, s, s, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last_3d).requires_grad_(True)
    w = (torch.randn(Cout, Cin, 3, 3, 3, device="cuda") * 0.02).contiguous(memory_format=torch.channels_last_3d).requires_grad_(True)
    b = torch.zeros(Cout, device="cuda")
    flops = 2.0 * N * s ** 3 * Cout * Cin * 27
    res = cb = None
    if epi:
        res = torch.randn(N, Cout, s, s, s, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last_3d)
        cb = torch.randn(N, Cout, device="cuda")
    y = ops.conv_nd(x, w, b, 1, 1, chan_bias=cb, residual=res)
    dy = torch.randn_like(y)
    y.backward(dy)
    torch.cuda.synchronize()
    ops.profile_start()
    for _ in range(reps):
        x.grad = None
        w.grad = None
        y = ops.conv_nd(x, w, b, 1, 1, chan_bias=cb, residual=res)
        y.backward(dy)
    torch.cuda.synchronize()
    prof = ops.profile_stop()
    agg = {}
    for kind, f, shape, a, e in prof:
        agg.setdefault(kind, []).append(a.elapsed_time(e))
    line = f"N={N} Cin={Cin} Cout={Cout} {s}^3 ({flops/1e9:.0f} GF): "
    for k in ("fwd", "dgrad", "wgrad"):
        ms = sorted(agg[k])[len(agg[k]) // 2]
        line += f"{k} {ms:.3f} ms {flops/ms/1e9:.0f} TF/s | "
    print(line, flush=True)
