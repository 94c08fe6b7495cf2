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
"""
from __future__ import annotations

import importlib
import sys
from typing import Dict, List, Tuple

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
    for modname, names in _DEFINING.items():
        try:
            mod = importlib.import_module(modname)
        except ImportError:
            if strict:
                raise
            continue
        for name in names:
            if hasattr(mod, name):
                originals[name] = getattr(mod, name)
                _saved.append((mod, name, originals[name]))
            setattr(mod, name, repl[name])
            patched.append(f"{modname}.{name}")
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


def uninstall() -> None:
    """Undo install()."""
    while _saved:
        mod, name, old = _saved.pop()
        setattr(mod, name, old)
