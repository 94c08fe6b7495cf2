"""Training attention (fwd + bwd) at the LDM shapes: fused flash kernels vs the unfused GEMM + softmax chain."""
import sys

import torch

sys.path.insert(0, ".")
from medical_image_generation_b200 import ops  # noqa: E402

torch.manual_seed(0)
for (B, L, C, heads) in ((8, 1728, 512, 1), (8, 216, 768, 1), (2, 6400, 512, 1), (2, 800, 768, 1), (1, 32768, 128, 1)):
    q, k, v = (torch.randn(B, L, C, device="cuda", dtype=torch.bfloat16).requires_grad_(True) for _ in range(3))
    dO = torch.randn(B, L, C, device="cuda", dtype=torch.bfloat16)
    scale = (C // heads) ** -0.5
    row = []
    for training_flash in (True, False):
        if not training_flash and B * heads * L * L * 4 > 8e9:
            row.append(float("nan"))
            continue
        ops.set_flash_attention(True, training=training_flash)
        def step():
            o = ops.sdpa(q, k, v, heads, scale)
            o.backward(dO)
        for _ in range(2):
            step()
        torch.cuda.synchronize()
        torch.cuda.reset_peak_memory_stats()
        base = torch.cuda.memory_allocated()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            step()
        e1.record()
        torch.cuda.synchronize()
        row.append(e0.elapsed_time(e1) / 5)
        row.append((torch.cuda.max_memory_allocated() - base) / 1e6)
    ops.set_flash_attention(True, training="auto")
    fl = 4.0 * B * L * L * C * 3.5   # fwd 2 GEMMs + bwd 5 GEMMs
    print(f"B={B} L={L} d={C // heads}: flash fwd+bwd {row[0]:.3f} ms ({fl / row[0] / 1e9:.0f} TFLOP/s algorithmic, peak extra memory "
          f"{row[1]:.0f} MB) | unfused {row[2] if len(row) > 2 else float('nan'):.3f} ms (peak extra memory {row[3] if len(row) > 3 else float('nan'):.0f} MB)",
          flush=True)
