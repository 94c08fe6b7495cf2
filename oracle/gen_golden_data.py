"""Golden vectors of the reference's data path, produced by the UNMODIFIED reference code.

    python -m oracle.gen_golden_data          -> tests/golden/data_path.pt

Runs `crop_and_pad_nd`, `MedicalDataset` (get_bbox, oversampling, configure_augmentation_params, __getitem__) and
`CustomBatchSampler` of /root/reference/medimgen/data_processing.py (AST-extracted, oracle/reference_loader.py) with
seeded `np.random`. Volumes are re-drawn from seeds (`make_case`), so the file holds only the small outputs."""
from __future__ import annotations

import hashlib
import os
import pickle
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tests", "golden", "data_path.pt")


def make_case(seed: int, shape):
    """A synthetic preprocessed case: fp32 (C, Z, Y, X) in [-0.2, 1.2] (so the clamp matters) and sampled class locations."""
    rs = np.random.RandomState(seed)
    image = (rs.rand(*shape) * 1.4 - 0.2).astype(np.float32)
    z, y, x = shape[1:]
    locs = {1: [(int(rs.randint(z)), int(rs.randint(y)), int(rs.randint(x))) for _ in range(20)],
            2: [] if seed % 2 else [(int(rs.randint(z)), int(rs.randint(y)), int(rs.randint(x))) for _ in range(5)]}
    return image, locs


TRANSFORM_ARGS = {"scaling": False, "rotation": False, "gaussian_noise": False, "gaussian_blur": False,
                  "low_resolution": False, "brightness": False, "contrast": False, "gamma": False, "mirror": False,
                  "dummy_2d": False}

DATASETS = {  # name -> (case shapes, patch_size, batch_size, oversample, channel_ids)
    "3d_fits": ([(2, 30, 42, 38), (2, 26, 50, 54), (2, 40, 34, 60)], (24, 32, 32), 4, 0.33, None),
    "3d_needs_padding": ([(1, 12, 20, 70), (1, 36, 18, 26), (1, 10, 50, 14)], (24, 36, 36), 3, 0.5, [0]),
    "3d_channel_select": ([(3, 24, 40, 40), (3, 30, 34, 52)], (16, 32, 32), 2, 1.0, [2, 0]),
    "2d": ([(1, 12, 70, 66), (1, 9, 50, 80), (1, 14, 64, 64)], (64, 64), 5, 0.4, None),
}


def main():
    from oracle import reference_loader as ref
    fn = ref.data_functions()
    crop, MedicalDataset, Sampler = fn["crop_and_pad_nd"], fn["MedicalDataset"], fn["CustomBatchSampler"]
    g = {"crop": [], "datasets": {}, "sampler": [], "aug_params": {}}

    # 1. crop_and_pad_nd (data_processing.py:150-225): inside, partly outside, wholly outside, fewer bbox dims than axes
    rs = np.random.RandomState(7)
    for shape, bbox, pad in [((2, 9, 10, 11), [[1, 5], [2, 9], [0, 11]], 0), ((2, 9, 10, 11), [[-3, 6], [4, 14], [-2, 13]], 0),
                             ((1, 9, 10, 11), [[9, 12], [0, 4], [0, 4]], 0), ((1, 9, 10, 11), [[-4, 0], [0, 4], [0, 4]], 0),
                             ((3, 8, 8), [[-2, 10], [3, 5]], 0), ((3, 8, 8), [[2, 6]], 0.5),
                             ((2, 6, 7, 8), [[-1, 7], [-1, 8], [-1, 9]], -1.0), ((1, 1, 12, 12), [[0, 1], [-2, 10], [5, 17]], 0)]:
        img = rs.rand(*shape).astype(np.float32)
        g["crop"].append({"seed_shape": shape, "bbox": bbox, "pad": pad, "image": torch.from_numpy(img),
                          "out": torch.from_numpy(crop(img, bbox, pad)),
                          "out_torch": crop(torch.from_numpy(img), bbox, pad)})

    # 2. MedicalDataset end to end with the third-party transformation pipeline = identity
    for name, (shapes, patch, bs, over, chans) in DATASETS.items():
        with tempfile.TemporaryDirectory() as d:
            ids = []
            for i, shp in enumerate(shapes):
                img, locs = make_case(100 + i, shp)
                np.save(os.path.join(d, f"case{i}.npy"), img)
                with open(os.path.join(d, f"case{i}.pkl"), "wb") as f:
                    pickle.dump({"class_locations": locs}, f)
                ids.append(f"case{i}")
            rec = {"shapes": shapes, "patch": patch, "batch_size": bs, "oversample": over, "channel_ids": chans,
                   "seeds": [100 + i for i in range(len(shapes))]}
            for section in ("training", "validation"):
                ds = MedicalDataset(d + "/", ids, bs, section, dict(TRANSFORM_ARGS, patch_size=list(patch)), over,
                                    channel_ids=chans)
                sampler = Sampler(ds, bs, number_of_steps=3, shuffle=section == "training")
                np.random.seed(2024)
                batches, bboxes, images = [], [], []
                for batch in sampler:
                    batches.append(batch)
                    for idx in batch:
                        # replay get_bbox on a copy of the RNG state to record the box, then the real __getitem__
                        state = np.random.get_state()
                        image, props = ds.load_image(ds.ids[idx[1]])
                        bboxes.append(ds.get_bbox(image.shape[1:], ds.oversampling_method(idx[0]), props["class_locations"],
                                                  is_2d=ds.patch_size[0] == 1))
                        np.random.set_state(state)
                        images.append(ds[idx]["image"])
                # bit-exact byte movement: step 0 is kept in full, every patch as a sha256 of its fp32 bytes
                rec[section] = {"batches": batches, "bboxes": bboxes, "images": torch.stack(images[:bs]),
                                "sha256": [hashlib.sha256(im.numpy().tobytes()).hexdigest() for im in images],
                                "shape": tuple(images[0].shape),
                                "initial_patch_size": tuple(ds.initial_patch_size), "need_to_pad": ds.need_to_pad.tolist()}
            g["datasets"][name] = rec

    # 3. CustomBatchSampler alone (data_processing.py:601-641): fewer cases than a batch, exact multiples, two epochs
    class _Len:
        def __init__(self, n): self.n = n
        def __len__(self): return self.n
    for n, bs, steps, shuffle in [(10, 4, 7, True), (3, 5, 4, True), (8, 4, 5, False), (1, 2, 3, True), (7, 7, 3, True)]:
        np.random.seed(99)
        s = Sampler(_Len(n), bs, number_of_steps=steps, shuffle=shuffle)
        g["sampler"].append({"n": n, "batch_size": bs, "steps": steps, "shuffle": shuffle,
                             "epochs": [list(s), list(s)]})

    # 4. soft augmentation parameters (data_processing.py:405-431)
    for patch in [(32, 40, 40), (64, 64)]:
        ds = MedicalDataset("/nonexistent/", [], 2, "training",
                            dict(TRANSFORM_ARGS, patch_size=list(patch), rotation=True, scaling=True, mirror=True,
                                 brightness=True, contrast=True, gamma=True), 0.0)
        ta = ds.transformation_args
        np.random.seed(5)
        dim = len(patch)
        rots = [[ta["rot_for_da"](None, a) for a in range(3)] for _ in range(4)]
        g["aug_params"][patch] = {"rot_draws": rots, "mirror_axes": ta["mirror_axes"], "scaling_range": ta["scaling_range"],
                                  "brightness_range": ta["brightness_range"], "contrast_range": ta["contrast_range"],
                                  "gamma_range": ta["gamma_range"], "dummy_2d": ta["dummy_2d"],
                                  "initial_patch_size": tuple(ds.initial_patch_size)}
    torch.save(g, OUT)
    print(f"data path goldens -> {OUT} ({os.path.getsize(OUT) / 1e3:.0f} kB)")


if __name__ == "__main__":
    main()
