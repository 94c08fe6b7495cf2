"""Import the UNMODIFIED reference model files from /root/reference (build container) or from the byte-for-byte copies
staged into the git-ignored baseline/_ref/ (GPU box; oracle/stage_reference.py) under the MONAI shim.

TEST / BENCH-BASELINE INFRASTRUCTURE. Used by
oracle/gen_golden.py to produce tests/golden/*.pt and by CPU tests (skipped when absent) that
pin oracle/torch_oracle.py against the real thing.
"""
from __future__ import annotations

import ast
import importlib.util
import os
import sys

_STAGED = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")
REFERENCE_ROOT = os.environ.get("MEDIMGEN_REFERENCE", "/root/reference")
if not os.path.isfile(os.path.join(REFERENCE_ROOT, "medimgen", "diffusion_model_unet_with_strides.py")):
    # the GPU box has no /root/reference: use the byte-for-byte copies staged by oracle/stage_reference.py (git-ignored)
    REFERENCE_ROOT = _STAGED
_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shim")
_cache: dict = {}


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "medimgen", "diffusion_model_unet_with_strides.py"))


def _load(name: str, filename: str):
    if name in _cache:
        return _cache[name]
    if not available():
        raise RuntimeError(f"reference not present at {REFERENCE_ROOT}")
    if _SHIM not in sys.path:
        sys.path.insert(0, _SHIM)
    spec = importlib.util.spec_from_file_location(name, os.path.join(REFERENCE_ROOT, "medimgen", filename))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    _cache[name] = mod
    return mod


def unet_module():
    """medimgen/diffusion_model_unet_with_strides.py as a module object."""
    return _load("_ref_medimgen_unet", "diffusion_model_unet_with_strides.py")


def ae_module():
    """medimgen/autoencoderkl_with_strides.py as a module object."""
    return _load("_ref_medimgen_ae", "autoencoderkl_with_strides.py")


def planner_functions() -> dict:
    """The four pure planner functions of medimgen/configuration.py:751-902, extracted by AST
    (the file itself imports nibabel/zarr/cv2 at top, which are absent)."""
    if "planner" in _cache:
        return _cache["planner"]
    import numpy as np
    path = os.path.join(REFERENCE_ROOT, "medimgen", "configuration.py")
    tree = ast.parse(open(path).read())
    want = {"compute_downsample_parameters", "compute_output_size", "create_autoencoder_dict", "create_ddpm_dict"}
    body = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in want]
    ns = {"np": np}
    exec(compile(ast.Module(body=body, type_ignores=[]), path, "exec"), ns)
    _cache["planner"] = {k: ns[k] for k in want}
    return _cache["planner"]


def rerandomize_zero_init(module, seed: int = 1234, std: float = 0.05):
    """zero_module() (unet:62-69) makes a fresh U-Net output exactly 0, so parity on fresh modules is
    vacuous. Re-draw every all-zero weight/bias tensor from N(0, std) with a fixed seed."""
    import torch
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for _, p in sorted(module.named_parameters()):
            if p.numel() and float(p.abs().max()) == 0.0:
                p.copy_(torch.randn(p.shape, generator=g) * std)
    return module
