"""Drop-in switch for the medimgen trainers.

`train_ldm.py`, `train_ddpm.py` and `train_autoencoder.py` bind their model / scheduler / inferer classes by name at
import time (`train_ldm.py:26-29`, `train_ddpm.py:17-19`, `train_autoencoder.py:22-27`):

    from medimgen.autoencoderkl_with_strides import AutoencoderKL
    from medimgen.diffusion_model_unet_with_strides import DiffusionModelUNet
    from generative.networks.schedulers import DDPMScheduler
    from generative.inferers import DiffusionInferer, LatentDiffusionInferer

`install()` rebinds exactly those names -- in the defining modules and in every trainer module that has already
imported them -- to the B200 implementations, so an unmodified trainer builds and drives the sm_100a path:

    import medical_image_generation_b200.compat as compat
    compat.install()
    from medimgen.train_ldm import LDM          # now constructs the B200 U-Net / scheduler / inferers

Nothing else of `generative` / `medimgen` is touched (VQVAE, discriminators, losses, metrics, data loading stay the
reference's). `uninstall()` restores the original bindings.

Where `monai-generative` is not installed at all, `install()` puts the package's own minimal `generative` on sys.path
(`medical_image_generation_b200/shims/generative`: scheduler + inferers = the B200 classes, placeholders for what the
step never executes). Where the reference's two model files cannot be imported (they need MONAI), modules of the same
names exposing the B200 classes are registered instead, so `from medimgen.diffusion_model_unet_with_strides import
DiffusionModelUNet` in an unmodified trainer still resolves.
"""
from __future__ import annotations

import importlib
import importlib.util
import os
import sys
import types
from typing import Dict, List, Tuple

_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")

_DEFINING = {
    "medimgen.diffusion_model_unet_with_strides": ("DiffusionModelUNet",),
    "medimgen.autoencoderkl_with_strides": ("AutoencoderKL",),
    "generative.networks.schedulers": ("DDPMScheduler",),
    "generative.inferers": ("DiffusionInferer", "LatentDiffusionInferer"),
}
_CONSUMERS = ("medimgen.train_ldm", "medimgen.train_ddpm", "medimgen.train_autoencoder")
_saved: List[Tuple[object, str, object]] = []


def _replacements() -> Dict[str, object]:
    from . import AutoencoderKL, DDPMScheduler, DiffusionInferer, DiffusionModelUNet, LatentDiffusionInferer
    return dict(AutoencoderKL=AutoencoderKL, DDPMScheduler=DDPMScheduler, DiffusionInferer=DiffusionInferer,
                DiffusionModelUNet=DiffusionModelUNet, LatentDiffusionInferer=LatentDiffusionInferer)


def install(strict: bool = False) -> List[str]:
    """Rebind the trainer-facing classes; returns the patched `module.name` strings. Modules that are not importable
    are skipped (or raise ImportError with strict=True)."""
    repl = _replacements()
    patched: List[str] = []
    originals = {}
    try:
        have_generative = importlib.util.find_spec("generative") is not None
    except (ImportError, ValueError):
        have_generative = "generative" in sys.modules
    if not have_generative and _SHIMS not in sys.path:
        sys.path.append(_SHIMS)          # appended: a real monai-generative always wins
        patched.append("sys.path+=shims/generative")
    for modname, names in _DEFINING.items():
        try:
            mod = importlib.import_module(modname)
        except ImportError:
            parent = modname.rsplit(".", 1)[0]
            try:
                importlib.import_module(parent)
            except ImportError:
                if strict:
                    raise
                continue
            # the package exists but this module's own imports fail (MONAI missing): stand-in module of the same name
            mod = types.ModuleType(modname)
            mod.__doc__ = f"registered by medical_image_generation_b200.compat: B200 classes under {modname}"
            sys.modules[modname] = mod
            _saved.append((sys.modules, modname, None))
        for name in names:
            if hasattr(mod, name):
                originals[name] = getattr(mod, name)
                _saved.append((mod, name, originals[name]))
            setattr(mod, name, repl[name])
            patched.append(f"{modname}.{name}")
    # the data pipeline (`from medimgen.data_processing import get_data_loaders`, train_ldm.py:34): where the reference's
    # module cannot be imported (zarr / blosc2 / batchgenerators(v2) missing) a module of the same name exposes the
    # resident-in-HBM loaders of .data; an importable reference module is left alone (opt in with
    # `compat.use_resident_loaders()`)
    if "medimgen.data_processing" not in sys.modules:
        try:
            importlib.import_module("medimgen.data_processing")
        except ImportError:
            try:
                importlib.import_module("medimgen")
                _register_data_module()
                patched.append("medimgen.data_processing (resident loaders)")
            except ImportError:
                if strict:
                    raise
    # trainer modules imported earlier hold their own references (`from x import Y`)
    for modname in _CONSUMERS:
        mod = sys.modules.get(modname)
        if mod is None:
            continue
        for name, new in repl.items():
            cur = getattr(mod, name, None)
            if cur is not None and cur is not new and (name not in originals or cur is originals[name]):
                _saved.append((mod, name, cur))
                setattr(mod, name, new)
                patched.append(f"{modname}.{name}")
    return patched


_DATA_NAMES = ("get_data_loaders", "create_split_files", "get_data_ids", "generate_crossval_split", "MedicalDataset",
               "CustomBatchSampler", "crop_and_pad_nd")


def _register_data_module() -> None:
    from . import data
    mod = types.ModuleType("medimgen.data_processing")
    mod.__doc__ = "registered by medical_image_generation_b200.compat: resident-in-HBM data path (data.py)"
    for name in _DATA_NAMES:
        setattr(mod, name, getattr(data, name))
    sys.modules["medimgen.data_processing"] = mod
    _saved.append((sys.modules, "medimgen.data_processing", None))


def use_resident_loaders() -> None:
    """Rebind `get_data_loaders` (and the dataset classes) of an importable `medimgen.data_processing`, and of trainer
    modules that already imported it, to the resident-in-HBM versions."""
    from . import data
    mod = importlib.import_module("medimgen.data_processing")
    for name in _DATA_NAMES:
        if hasattr(mod, name):
            _saved.append((mod, name, getattr(mod, name)))
        setattr(mod, name, getattr(data, name))
    for modname in _CONSUMERS:
        tm = sys.modules.get(modname)
        if tm is not None and hasattr(tm, "get_data_loaders"):
            _saved.append((tm, "get_data_loaders", tm.get_data_loaders))
            tm.get_data_loaders = data.get_data_loaders


def uninstall() -> None:
    """Undo install()."""
    while _saved:
        mod, name, old = _saved.pop()
        if mod is sys.modules:
            sys.modules.pop(name, None)
        else:
            setattr(mod, name, old)
    if _SHIMS in sys.path:
        sys.path.remove(_SHIMS)
        for k in [k for k in sys.modules if k == "generative" or k.startswith("generative.")]:
            if getattr(sys.modules[k], "__file__", None) and sys.modules[k].__file__.startswith(_SHIMS):
                del sys.modules[k]
