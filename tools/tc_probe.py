"""Diagnostic for the tcgen05 engine on a real B200: runs a ladder of conv shapes through engine=2 (tcgen05) and
engine=1 (SIMT) and prints relative errors against torch fp32 on the CPU. Never raises; prints one line per case."""
import math
import sys
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, ".")
from medical_image_generation_b200 import ops  # noqa: E402


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def run(N, Cin, Cout, sp, k, s, p, what="fwd"):
    g = torch.Generator().manual_seed(1)
    x = torch.randn((N, Cin, *sp), generator=g).bfloat16().float()
    w = (torch.randn((Cout, Cin, *k), generator=g) / math.sqrt(Cin * math.prod(k))).bfloat16().float()
    b = torch.randn(Cout, generator=g) * 0.1
    xr = x.clone().requires_grad_(True)
    y_ref = F.conv3d(xr, w, b, stride=s, padding=p)
    probe = torch.randn(y_ref.shape, generator=g).bfloat16().float()
    (y_ref * probe).sum().backward()
    out = {}
    for eng in (1, 2):
        ops.set_engine(eng)
        try:
            xd = x.cuda().bfloat16().contiguous(memory_format=torch.channels_last_3d).requires_grad_(True)
            wd = w.cuda().contiguous(memory_format=torch.channels_last_3d).requires_grad_(True)
            torch.cuda.synchronize()
            t0 = time.time()
            y = ops.conv_nd(xd, wd, b.cuda(), s, p)
            torch.cuda.synchronize()
            t1 = time.time()
            ops.set_engine(0 if eng == 2 else 1)  # wgrad has no tcgen05 kernel yet: auto for backward
            y.backward(probe.cuda().bfloat16().contiguous(memory_format=torch.channels_last_3d))
            torch.cuda.synchronize()
            out[eng] = (rel(y, y_ref), rel(xd.grad, xr.grad), (t1 - t0) * 1e3)
        except Exception as e:  # noqa: BLE001
            out[eng] = ("ERR " + str(e)[:120],)
        finally:
            ops.set_engine(0)
    print(f"N={N} Cin={Cin} Cout={Cout} sp={sp} k={k} s={s} p={p}: simt={out[1]} tc={out[2]}", flush=True)


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), flush=True)
    run(1, 64, 64, (4, 4, 8), (1, 1, 1), (1, 1, 1), (0, 0, 0))      # one tile, K=64: plain GEMM
    run(1, 64, 32, (4, 4, 8), (1, 1, 1), (1, 1, 1), (0, 0, 0))
    run(1, 128, 64, (4, 4, 8), (1, 1, 1), (1, 1, 1), (0, 0, 0))     # two K blocks
    run(1, 64, 256, (4, 8, 8), (1, 1, 1), (1, 1, 1), (0, 0, 0))     # BN=256
    run(1, 64, 64, (4, 4, 8), (3, 3, 3), (1, 1, 1), (1, 1, 1))      # taps + halo
    run(2, 32, 32, (5, 6, 7), (3, 3, 3), (1, 1, 1), (1, 1, 1))      # Cin=32: two taps per K block, M tail
    run(2, 32, 32, (9, 8, 7), (3, 3, 3), (2, 2, 2), (1, 1, 1))      # strided (dgrad with holes)
    run(1, 256, 256, (12, 12, 12), (3, 3, 3), (1, 1, 1), (1, 1, 1))  # realistic level-1 shape
    run(8, 768, 768, (6, 6, 6), (3, 3, 3), (1, 1, 1), (1, 1, 1))     # deep level: split-K path
    run(1, 256, 8, (8, 8, 8), (3, 3, 3), (1, 1, 1), (1, 1, 1))       # narrow N
