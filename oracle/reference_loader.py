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


def data_functions() -> dict:
    """`crop_and_pad_nd`, `MedicalDataset` and `CustomBatchSampler` of medimgen/data_processing.py:150-225,274-641, extracted by
    AST and executed as written (the file itself imports zarr / blosc2 / batchgenerators(v2) at top, all absent here).
    The names the extracted code touches are bound to the real numpy / torch objects; `zarr` / `blosc2` are empty stand-ins
    (only used in isinstance checks and for file formats that cannot be read in this image), and
    `define_nnunet_transformations` -- the batchgeneratorsv2 pipeline, a third-party dependency -- is the identity, so
    `MedicalDataset.__getitem__` yields crop + pad + channel selection + clamp exactly as the reference computes them."""
    if "data" in _cache:
        return _cache["data"]
    import glob
    import pickle
    import types
    from functools import partial
    from typing import List, Tuple, Union

    import numpy as np
    import torch
    import torch.nn.functional as F
    from torch.utils.data import Dataset, Sampler

    path = os.path.join(REFERENCE_ROOT, "medimgen", "data_processing.py")
    tree = ast.parse(open(path).read())
    want = {"crop_and_pad_nd", "MedicalDataset", "CustomBatchSampler", "generate_crossval_split", "create_split_files",
            "get_data_ids"}
    body = [n for n in tree.body if isinstance(n, (ast.FunctionDef, ast.ClassDef)) and n.name in want]

    class _Never:   # isinstance(x, _Never) is always False for arrays
        pass
    blosc2 = types.SimpleNamespace(ndarray=types.SimpleNamespace(NDArray=_Never), open=None)
    zarr = types.SimpleNamespace(core=types.SimpleNamespace(Array=_Never), open_group=None)

    def identity_pipeline(params, validation=False):
        return lambda image: {"image": image}

    import json
    from sklearn.model_selection import KFold, train_test_split
    ns = {"json": json, "KFold": KFold, "train_test_split": train_test_split,
          "np": np, "torch": torch, "F": F, "os": os, "glob": glob, "pickle": pickle, "partial": partial,
          "Dataset": Dataset, "Sampler": Sampler, "List": List, "Tuple": Tuple, "Union": Union, "blosc2": blosc2,
          "zarr": zarr, "define_nnunet_transformations": identity_pipeline}
    exec(compile(ast.Module(body=body, type_ignores=[]), path, "exec"), ns)
    _cache["data"] = {k: ns[k] for k in want}
    return _cache["data"]


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
