"""tcgen05 GEMM throughput for the four operand-major combinations (is MN-major slower than K-major?)."""
import sys

import torch

sys.path.insert(0, ".")
from medical_image_generation_b200 import ops  # noqa: E402

M = N = 4096
K = 8192
print(torch.cuda.get_device_name(0))
for a_mn in (0, 1):
    for b_mn in (0, 1):
        A = torch.randn((K, M) if a_mn else (M, K), device="cuda").bfloat16()
        Bm = torch.randn((K, N) if b_mn else (N, K), device="cuda").bfloat16()
        Cm = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        a = (1, M, 0, 0) if a_mn else (K, 1, 0, 0)      # (a_m, a_k, outer, inner)
        b = (N, 1, 0, 0) if b_mn else (1, K, 0, 0)      # (b_k, b_n, outer, inner)
        for _ in range(3):
            ops._gemm(A, Bm, Cm, M, N, K, 1, 1, a, b, (N, 1, 0, 0))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            ops._gemm(A, Bm, Cm, M, N, K, 1, 1, a, b, (N, 1, 0, 0))
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        ref = (A.float().t() if a_mn else A.float()) @ (Bm.float() if b_mn else Bm.float().t())
        err = float((Cm.float() - ref).norm() / ref.norm())
        print(f"A {'MN' if a_mn else 'K '}-major, B {'MN' if b_mn else 'K '}-major: {ms:.3f} ms  {2.0*M*N*K/ms/1e9:.0f} TF/s  rel err {err:.2e}", flush=True)
