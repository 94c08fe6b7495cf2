"""CPU restatement (numpy) of the reference's data path: `MedicalDataset.__getitem__` and its helpers,
medimgen/data_processing.py:150-225 (crop_and_pad_nd), :405-431 (soft augmentation parameters), :433-441 (foreground
oversampling), :463-527 (get_bbox), :558-598 (__getitem__), :601-641 (CustomBatchSampler).

TEST INFRASTRUCTURE ONLY: imported by tests/, never by the product package (medical_image_generation_b200/data.py runs
these steps on the GPU through the C ABI and raises without the CUDA extension).

Pinned: `oracle/gen_golden_data.py` executes the UNMODIFIED reference functions (AST-extracted, see
oracle/reference_loader.data_functions) with seeded `np.random` and commits inputs/outputs to tests/golden/data_path.pt;
tests/test_oracle_golden.py checks every function here against them.

PARITY UNPINNED for the augmentation transforms below the line "batchgeneratorsv2": that package is a third-party
dependency (pyproject.toml, unpinned) absent from /root/reference and from this image; the restatement follows its
published transforms as the reference configures them (data_processing.py:745-858) and is anchored on torch's own
`grid_sample`, `flip`, `mean`, `std`, `pow`."""
from __future__ import annotations

import numpy as np


# ------------------------------------------------------------------------------------------ pinned to the reference
def crop_and_pad_nd(image: np.ndarray, bbox, pad_value=0) -> np.ndarray:
    """data_processing.py:150-225. bbox covers the LAST len(bbox) dims, upper bound excluded."""
    nd, cd = image.ndim, len(bbox)
    target = list(image.shape[:nd - cd]) + [hi - lo for lo, hi in bbox]
    slices, pads = [slice(None)] * (nd - cd), [(0, 0)] * (nd - cd)
    for (lo, hi), size in zip(bbox, image.shape[nd - cd:]):
        if hi <= 0 or lo >= size:            # :195-201: entirely outside -> zeros (NOT pad_value, as the reference)
            return np.zeros(target, dtype=image.dtype)
        slices.append(slice(max(lo, 0), min(hi, size)))
        pads.append((max(0, -lo), max(0, hi - size)))
    return np.pad(image[tuple(slices)], pads, mode="constant", constant_values=pad_value)


def oversample_last_xx_percent(sample_idx: int, batch_size: int, oversample_foreground_percent: float) -> bool:
    """data_processing.py:433-436 (Python round: banker's rounding)."""
    return sample_idx >= round(batch_size * (1 - oversample_foreground_percent))


def get_bbox(data_shape, force_fg, class_locations, initial_patch_size, need_to_pad, is_2d=False, rng=np.random):
    """data_processing.py:463-527; consumes `rng` in the reference's order."""
    dim = len(data_shape)
    need_to_pad = np.array(need_to_pad).copy()
    for d in range(dim):
        if need_to_pad[d] + data_shape[d] < initial_patch_size[d]:
            need_to_pad[d] = initial_patch_size[d] - data_shape[d]
    lbs = [-need_to_pad[i] // 2 for i in range(dim)]
    ubs = [data_shape[i] + need_to_pad[i] // 2 + need_to_pad[i] % 2 - initial_patch_size[i] for i in range(dim)]
    bbox_lbs = [rng.randint(lbs[i], ubs[i] + 1) for i in range(dim)]
    if force_fg and class_locations is not None:
        eligible = [c for c in class_locations if len(class_locations[c]) > 0]
        if eligible:
            cls = rng.choice(eligible)
            voxels = class_locations[cls]
            voxel = voxels[rng.choice(len(voxels))]
            for i in range(dim):
                if is_2d and i == 0:
                    bbox_lbs[0] = voxel[0]
                elif not is_2d:
                    bbox_lbs[i] = max(lbs[i], min(voxel[i] - initial_patch_size[i] // 2, ubs[i]))
    for i in range(dim - 2, dim):            # :507-524: (jittered) centre crop in the last two axes
        crop, size = initial_patch_size[i], data_shape[i]
        center = size // 2
        if size < crop:
            bbox_lbs[i] = center - crop // 2
        else:
            max_offset = min(10, center - crop // 2, size - center - (crop - crop // 2))
            offset = rng.randint(-max_offset, max_offset + 1) if max_offset > 0 else 0
            bbox_lbs[i] = center + offset - crop // 2
    return bbox_lbs, [bbox_lbs[i] + initial_patch_size[i] for i in range(dim)]


def sampler_batches(n_cases: int, batch_size: int, number_of_steps: int, shuffle: bool, rng=np.random):
    """CustomBatchSampler.define_indices + __iter__, data_processing.py:610-638: list of batches of
    (position in batch, case index). `shuffle` permutes the sampler's own index list in place (state kept by the caller:
    pass `indices` back in for the next epoch through the return value)."""
    indices = list(range(n_cases))
    return sampler_batches_from(indices, batch_size, number_of_steps, shuffle, rng)


def sampler_batches_from(indices: list, batch_size: int, number_of_steps: int, shuffle: bool, rng=np.random):
    if shuffle:
        rng.shuffle(indices)
    order, available = [], indices.copy()
    while len(order) < number_of_steps * batch_size:
        if len(available) < batch_size:
            available = indices.copy()
            if shuffle:
                rng.shuffle(available)
        order.extend(available[:batch_size])
        available = available[batch_size:]
    return [[(i, s) for i, s in enumerate(order[k * batch_size:(k + 1) * batch_size])] for k in range(number_of_steps)]


def soft_augmentation_params(patch_size) -> dict:
    """configure_augmentation_params(heavy_augmentation=False), data_processing.py:405-431. `rot_for_da` is returned as
    (axis, low, high): the reference's closure draws uniform(-0.174533, 0.174533) for that axis and 0 for the others."""
    dim = len(patch_size)
    return {"rot_for_da": (0 if dim == 3 else 2 if dim == 2 else None, -0.174533, 0.174533), "do_dummy_2d": False,
            "initial_patch_size": tuple(patch_size), "mirror_axes": (2,) if dim == 3 else (1,),
            "scale_range": (0.9, 1.1), "brightness_range": (0.9, 1.1), "contrast_range": (0.9, 1.1),
            "gamma_range": (0.9, 1.1)}


def getitem_untransformed(image: np.ndarray, class_locations, batch_idx: int, batch_size: int, patch_size,
                          oversample_foreground_percent: float, channel_ids=None, rng=np.random) -> np.ndarray:
    """MedicalDataset.__getitem__ with the transformation pipeline = identity, data_processing.py:558-595."""
    patch = (1, *patch_size) if len(patch_size) == 2 else tuple(patch_size)
    force_fg = oversample_last_xx_percent(batch_idx, batch_size, oversample_foreground_percent)
    lbs, ubs = get_bbox(image.shape[1:], force_fg, class_locations, patch, np.zeros(len(patch), dtype=int),
                        is_2d=patch[0] == 1, rng=rng)
    out = crop_and_pad_nd(image, [[a, b] for a, b in zip(lbs, ubs)], 0)
    if channel_ids is not None:
        out = out[channel_ids, ...]
    if patch[0] == 1:
        out = np.squeeze(out, axis=1)
    return np.clip(out.astype(np.float32), 0.0, 1.0)


# ------------------------------------------------------------------ batchgeneratorsv2 [upstream-memory, parity unpinned]
def brightness(img, multipliers):
    """MultiplicativeBrightnessTransform: img[c] *= m_c."""
    out = np.array(img, dtype=np.float32, copy=True)
    for c, m in enumerate(multipliers):
        out[c] *= np.float32(m)
    return out


def contrast(img, multipliers):
    """ContrastTransform(preserve_range=True): (x - mean) * m + mean, clamped to the channel's former [min, max]."""
    import torch
    out = torch.as_tensor(np.array(img, dtype=np.float32, copy=True))
    for c, m in enumerate(multipliers):
        if m is None:
            continue
        mean, lo, hi = out[c].mean(), out[c].min(), out[c].max()
        out[c] = ((out[c] - mean) * float(m) + mean).clamp(lo, hi)
    return out.numpy()


def gamma(img, gammas, invert=False, retain_stats=True):
    """GammaTransform: pow((x - min) / max(range, 1e-7), g) * range + min on (optionally negated) x; with retain_stats
    the channel is re-standardised to its former mean / (unbiased) std."""
    import torch
    out = torch.as_tensor(np.array(img, dtype=np.float32, copy=True))
    for c, g in enumerate(gammas):
        if g is None:
            continue
        x = -out[c] if invert else out[c]
        if retain_stats:
            mean, std = x.mean(), x.std()
        lo = x.min()
        rnge = x.max() - lo
        x = torch.pow((x - lo) / rnge.clamp(min=1e-7), float(g)) * rnge + lo
        if retain_stats:
            x = (x - x.mean()) * (std / x.std().clamp(min=1e-7)) + mean
        out[c] = -x if invert else x
    return out.numpy()


def mirror(img, axes):
    """MirrorTransform: flip the listed SPATIAL axes (axis 0 = first axis after the channel)."""
    return np.flip(img, [a + 1 for a in axes]).copy() if len(axes) else np.array(img, copy=True)


def rotation_scale_matrix(angles, scales) -> np.ndarray:
    """3x3 matrix M in (z, y, x) order: source offset from the patch centre = M @ output offset.
    angles = rotation about the z, y, x axes (the reference's soft augmentation only rotates about axis 0)."""
    az, ay, ax = angles
    cz, sz, cy, sy, cx, sx = np.cos(az), np.sin(az), np.cos(ay), np.sin(ay), np.cos(ax), np.sin(ax)
    rz = np.array([[1, 0, 0], [0, cz, -sz], [0, sz, cz]])      # about axis 0 (z): mixes y and x
    ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])      # about axis 1 (y): mixes z and x
    rx = np.array([[cx, -sx, 0], [sx, cx, 0], [0, 0, 1]])      # about axis 2 (x): mixes z and y
    return (rz @ ry @ rx @ np.diag(np.asarray(scales, dtype=np.float64))).astype(np.float32)


def affine_resample(patch: np.ndarray, mat: np.ndarray) -> np.ndarray:
    """SpatialTransform's resampling step: trilinear `grid_sample(..., mode='bilinear', padding_mode='zeros',
    align_corners=False)` of the cropped patch (C, Z, Y, X) at centre + M @ offset."""
    import torch
    import torch.nn.functional as F
    x = torch.as_tensor(np.ascontiguousarray(patch), dtype=torch.float32)[None]
    size = np.array(x.shape[2:], dtype=np.float64)
    ctr = (size - 1) / 2
    zz, yy, xx = np.meshgrid(*[np.arange(s) - c for s, c in zip(size.astype(int), ctr)], indexing="ij")
    off = np.stack([zz, yy, xx], -1).astype(np.float32)                      # (Z, Y, X, 3) in z, y, x
    pos = off @ np.asarray(mat, dtype=np.float32).T + ctr.astype(np.float32)  # source voxel coordinates
    norm = (2 * pos + 1) / size.astype(np.float32) - 1                       # align_corners=False
    grid = torch.as_tensor(norm[..., ::-1].copy(), dtype=torch.float32)[None]  # grid_sample wants (x, y, z)
    return F.grid_sample(x, grid, mode="bilinear", padding_mode="zeros", align_corners=False)[0].numpy()
