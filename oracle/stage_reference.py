"""Stage the UNMODIFIED reference sources into the git-ignored `baseline/_ref/` so they travel to the GPU box.

    python -m oracle.stage_reference          (also run by __graft_entry__.build() when /root/reference is present)

`/root/reference` does not exist on the GPU box. `baseline/_ref/` is git-ignored (never committed: reference sources
stay out of the history) but NOT gpurun-ignored, so `bench.py --impl reference` can time the real reference modules
there and the drop-in test can drive the real `train_ldm.py`. Files are byte-for-byte copies (sha256 recorded in
`baseline/_ref/MANIFEST.json`); nothing here is imported by the product package."""
from __future__ import annotations

import hashlib
import json
import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("MEDIMGEN_REFERENCE", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")
FILES = ["__init__.py", "diffusion_model_unet_with_strides.py", "autoencoderkl_with_strides.py", "train_ldm.py",
         "train_autoencoder.py", "train_ddpm.py", "utils.py", "configuration.py", "data_processing.py"]


def stage(verbose: bool = True) -> str | None:
    src_pkg = os.path.join(SRC, "medimgen")
    if not os.path.isdir(src_pkg):
        return DST if os.path.isfile(os.path.join(DST, "MANIFEST.json")) else None
    dst_pkg = os.path.join(DST, "medimgen")
    os.makedirs(dst_pkg, exist_ok=True)
    manifest = {}
    for name in FILES:
        s = os.path.join(src_pkg, name)
        if not os.path.isfile(s):
            continue
        shutil.copyfile(s, os.path.join(dst_pkg, name))
        manifest[name] = hashlib.sha256(open(s, "rb").read()).hexdigest()
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": src_pkg, "sha256": manifest}, f, indent=1)
    if verbose:
        print(f"staged {len(manifest)} unmodified reference files into {dst_pkg}")
    return DST


if __name__ == "__main__":
    stage()
