"""Shared helpers for the parity tests (TEST CODE: may import oracle/)."""
import torch

FP32_TOL = 1e-4   # north_star: per-block fp32 outputs within 1e-4 relative error
BF16_TOL = 2e-2   # north_star: bf16 within 2e-2


def rel_err(got, want, floor=1e-30):
    got, want = got.detach().double().cpu(), want.detach().double().cpu()
    return float((got - want).norm() / want.norm().clamp_min(floor))


def bf16_round(t):
    return t.to(torch.bfloat16).to(torch.float32)


def tol_for(dtype):
    return FP32_TOL if dtype == torch.float32 else BF16_TOL
