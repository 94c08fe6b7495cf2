"""Data path of `medimgen` on the B200: the preprocessed cases live in HBM and a training batch
is cut, augmented and clamped by CUDA kernels (csrc/data_path.cu, C ABI `mig_patch_*`) instead of by DataLoader workers.

Mirrors medimgen/data_processing.py:
  * `MedicalDataset(data_path, data_ids, batch_size, section, transformation_args, oversample_foreground_percent,
    channel_ids=None, probabilistic_oversampling=False)` (:274-598) -- same constructor, `get_bbox`, oversampling rules and
    `__getitem__((batch_idx, sample_idx)) -> {'id', 'image'}`; the sampling decisions consume `np.random` in the
    reference's order, so a seeded run cuts the same boxes (tests/test_data_oracle.py, tests/golden/data_path.pt);
  * `CustomBatchSampler` (:601-641), `crop_and_pad_nd` (:150-225), `create_split_files` / `get_data_ids` /
    `get_data_loaders` (:34-147);
  * readers for what `load_image` (:536-556) accepts: `.zarr` (v2 directory store, read without the zarr package), `.npy`,
    `.npz`; Blosc-compressed chunks (what the reference's preprocessing writes, configuration.py:1404) need `numcodecs`
    and `.b2nd` needs `blosc2` -- neither is in this image, the readers say so instead of guessing.

The augmentation pipeline (data_processing.py:745-858) is batchgeneratorsv2, a third-party package that is not in this
image [upstream-memory, parity unpinned]: the transforms the planner switches on (configuration.py:933-945: scaling,
rotation about the slice axis, brightness, contrast, gamma, mirror) are restated with the probabilities and ranges the
reference passes; gaussian noise / blur / low-resolution simulation (off in every planner output) raise.

There is no CPU path: the cases are uploaded once and every batch is produced by the kernels; a missing shared library
or device raises.
"""
from __future__ import annotations

import ctypes as C
import glob
import json
import math
import os
import pickle
from typing import Optional

import numpy as np
import torch

from . import _lib

PATCH_MAX_CH = 8
_DESC = np.dtype([("src_offset", "<i8"), ("src_dims", "<i4", 4), ("lb", "<i4", 3), ("flip", "<i4", 3),
                  ("affine", "<i4"), ("mat", "<f4", 9), ("mult", "<f4", PATCH_MAX_CH), ("channel", "<i4", PATCH_MAX_CH)])
assert _DESC.itemsize == 152   # mig_patch_desc (include/medimgen_b200.h)
_I3 = C.c_int32 * 3
NO_CLAMP = (1.0, 0.0)          # lo > hi


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t: torch.Tensor):
    return C.c_void_p(t.data_ptr())


def _dtype_code(dtype) -> int:
    if dtype == torch.float32:
        return _lib.F32
    if dtype == torch.bfloat16:
        return _lib.BF16
    raise ValueError(f"data path output dtype must be float32 or bfloat16, got {dtype}")


# ------------------------------------------------------------------------------------------------------ file readers
def _decompressor(spec):
    if spec is None:
        return lambda b: b
    cid = spec.get("id")
    if cid in ("zlib", "gzip"):
        import zlib
        return lambda b: zlib.decompress(b, 15 + 32)
    if cid == "bz2":
        import bz2
        return bz2.decompress
    if cid == "lzma":
        import lzma
        return lzma.decompress
    if cid == "blosc":
        try:
            import numcodecs
            codec = numcodecs.get_codec(spec)
            return codec.decode
        except ImportError as e:
            raise RuntimeError("zarr chunk is Blosc-compressed (what medimgen's preprocessing writes, "
                               "configuration.py:1404) and `numcodecs` is not installed; install it or re-save the "
                               "dataset with a zlib / null compressor") from e
    raise RuntimeError(f"unsupported zarr compressor {spec!r}")


def read_zarr_array(path: str) -> np.ndarray:
    """Read a whole zarr v2 array directory (`.zarray` + chunk files) into memory."""
    with open(os.path.join(path, ".zarray")) as f:
        meta = json.load(f)
    if meta.get("zarr_format") != 2:
        raise RuntimeError(f"{path}: only zarr format 2 is read here (got {meta.get('zarr_format')})")
    if meta.get("filters"):
        raise RuntimeError(f"{path}: zarr filters are not supported")
    shape, chunks = tuple(meta["shape"]), tuple(meta["chunks"])
    dtype, order = np.dtype(meta["dtype"]), meta.get("order", "C")
    sep = meta.get("dimension_separator", ".")
    fill = meta.get("fill_value")
    out = np.empty(shape, dtype=dtype)
    decode = _decompressor(meta.get("compressor"))
    grid = [math.ceil(s / c) for s, c in zip(shape, chunks)]
    for idx in np.ndindex(*grid):
        name = os.path.join(path, sep.join(str(i) for i in idx)) if idx else os.path.join(path, "0")
        sel = tuple(slice(i * c, min((i + 1) * c, s)) for i, c, s in zip(idx, chunks, shape))
        if not os.path.isfile(name):
            out[sel] = 0 if fill is None else fill
            continue
        with open(name, "rb") as f:
            raw = decode(f.read())
        chunk = np.frombuffer(raw, dtype=dtype).reshape(chunks, order=order)
        out[sel] = chunk[tuple(slice(0, s.stop - s.start) for s in sel)]
    return out


def load_case(data_path: str, name: str):
    """`MedicalDataset.load_image`, data_processing.py:536-556: (image (C, Z, Y, X), properties)."""
    zarr_path = os.path.join(data_path, name + ".zarr")
    if os.path.isdir(zarr_path):
        image = read_zarr_array(os.path.join(zarr_path, "image"))
    elif os.path.isfile(os.path.join(data_path, name + ".npy")):
        image = np.load(os.path.join(data_path, name + ".npy"), mmap_mode="r")
    elif os.path.isfile(os.path.join(data_path, name + ".npz")):
        image = np.load(os.path.join(data_path, name + ".npz"))["data"]
    elif os.path.isfile(os.path.join(data_path, name + ".b2nd")):
        raise RuntimeError(f"{name}.b2nd needs the `blosc2` package, which is not installed")
    else:
        raise FileNotFoundError(f"no .zarr / .npy / .npz / .b2nd for case {name!r} under {data_path}")
    with open(os.path.join(data_path, name + ".pkl"), "rb") as f:
        properties = pickle.load(f)
    return image, properties


# -------------------------------------------------------------------------------------------------- resident volumes
class ResidentVolumes:
    """All cases of a split as ONE fp32 device buffer (a 160x160x128 channel is 13 MB; thousands of cases fit 180 GB of HBM3e)."""

    def __init__(self, device="cuda"):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("medical_image_generation_b200.data keeps the dataset in GPU memory; device must be CUDA")
        _lib.load()
        self.names: list[str] = []
        self.shapes: list[tuple] = []
        self.offsets: list[int] = []
        self._pending: list[np.ndarray] = []
        self.buffer: Optional[torch.Tensor] = None

    def add(self, name: str, image) -> int:
        """image: (C, Z, Y, X) numpy array (any dtype, stored as fp32) or an fp32 tensor already on the device."""
        if not isinstance(image, torch.Tensor):
            image = np.asarray(image)
        if image.ndim != 4:
            raise ValueError(f"case {name!r}: expected (C, Z, Y, X), got shape {tuple(image.shape)}")
        self.names.append(name)
        self.shapes.append(tuple(int(v) for v in image.shape))
        self._pending.append(image)
        return len(self.names) - 1

    def finalize(self):
        sizes = [int(np.prod(s)) for s in self.shapes]
        self.offsets = [0]
        for s in sizes[:-1]:
            self.offsets.append(self.offsets[-1] + (s + 63) // 64 * 64)   # 256-byte aligned cases
        total = (self.offsets[-1] + sizes[-1]) if sizes else 0
        self.buffer = torch.zeros(max(total, 1), dtype=torch.float32, device=self.device)
        for img, off, n in zip(self._pending, self.offsets, sizes):
            if isinstance(img, torch.Tensor):
                src = img.to(dtype=torch.float32).reshape(-1)
            else:
                src = torch.from_numpy(np.ascontiguousarray(img, dtype=np.float32)).reshape(-1)
            self.buffer[off:off + n].copy_(src)
        self._pending = []
        return self


# -------------------------------------------------------------------------------------------------------- kernels
class _Scratch:
    """Pinned host + device twins for the small per-batch parameter tables. The device copy is consumed in stream order;
    the pinned copy is only rewritten after the previous upload has left it (event)."""

    def __init__(self, device):
        self.device = device
        self._bufs: dict = {}

    def upload(self, key: str, host: np.ndarray) -> torch.Tensor:
        raw = np.ascontiguousarray(host).view(np.uint8).reshape(-1)
        ent = self._bufs.get(key)
        if ent is None or ent[0].numel() < raw.size:
            cap = (max(raw.size, 256) + 255) // 256 * 256
            ent = [torch.empty(cap, dtype=torch.uint8).pin_memory(),
                   torch.empty(cap, dtype=torch.uint8, device=self.device), torch.cuda.Event()]
            self._bufs[key] = ent
        else:
            ent[2].synchronize()
        pin, dev, event = ent
        pin[:raw.size].copy_(torch.from_numpy(raw))
        dev[:raw.size].copy_(pin[:raw.size], non_blocking=True)
        event.record()
        return dev


def _on_one_device(what: str, *tensors) -> torch.device:
    dev = tensors[0].device
    for t in tensors:
        if t is not None and (not t.is_cuda or t.device != dev):
            raise RuntimeError(f"{what}: every tensor must live on the same CUDA device (got {t.device} and {dev}); "
                               "this package has no CPU path")
    return dev


def patch_gather(volumes: torch.Tensor, descs_dev: torch.Tensor, out: torch.Tensor, B: int, C_: int, patch, *,
                 channels_last: bool = False, any_affine: bool = False, pad_value: float = 0.0, clamp=NO_CLAMP):
    """`mig_patch_gather`: B patches of C_ channels cut out of `volumes` as `descs_dev` (B packed `mig_patch_desc`) says."""
    if volumes.dtype != torch.float32:
        raise RuntimeError("patch_gather: the resident cases are float32")
    with torch.cuda.device(_on_one_device("patch_gather", volumes, descs_dev, out)):
        _lib.call("mig_patch_gather", _ptr(volumes), _ptr(descs_dev), _ptr(out), _dtype_code(out.dtype), B, C_,
                  _I3(*[int(v) for v in patch]), int(channels_last), int(any_affine), float(pad_value), float(clamp[0]),
                  float(clamp[1]), _stream())
    return out


def patch_stats(x: torch.Tensor, rows: int, S: int, active: Optional[torch.Tensor] = None,
                out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """{mean, unbiased std, min, max} per row of S fp32 voxels."""
    if x.dtype != torch.float32 or not x.is_cuda:
        raise RuntimeError("patch_stats: needs a CUDA float32 tensor")
    lib = _lib.load()
    if out is None:
        out = torch.zeros(rows, 4, dtype=torch.float32, device=x.device)
    ws_bytes = int(lib.mig_patch_stats_workspace_bytes(rows))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
    with torch.cuda.device(_on_one_device("patch_stats", x, out, active)):
        _lib.call("mig_patch_stats", _ptr(x), _ptr(out), None if active is None else _ptr(active), rows, S, _ptr(ws),
                  ws_bytes, _stream())
    return out


def patch_intensity(x: torch.Tensor, y: torch.Tensor, op: torch.Tensor, stats0: torch.Tensor, stats1: torch.Tensor,
                    B: int, C_: int, S: int, *, channels_last: bool = False, clamp=NO_CLAMP):
    """`mig_patch_intensity`: per-row contrast / gamma / re-standardisation (op = rows x {mode, param, invert, 0})."""
    if x.dtype != torch.float32:
        raise RuntimeError("patch_intensity: the staged batch is float32")
    with torch.cuda.device(_on_one_device("patch_intensity", x, y, op, stats0, stats1)):
        _lib.call("mig_patch_intensity", _ptr(x), _ptr(y), _dtype_code(y.dtype), _ptr(op), _ptr(stats0), _ptr(stats1), B,
                  C_, S, int(channels_last), float(clamp[0]), float(clamp[1]), _stream())
    return y


def crop_and_pad_nd(image: torch.Tensor, bbox, pad_value=0) -> torch.Tensor:
    """data_processing.py:150-225 for a CUDA float32 tensor: crop the LAST len(bbox) (<= 3) axes to
    [lo, hi) per axis, padding what lies outside the tensor with `pad_value` (a box entirely outside returns zeros, as the
    reference does)."""
    if not isinstance(image, torch.Tensor) or not image.is_cuda or image.dtype != torch.float32:
        raise RuntimeError("crop_and_pad_nd: needs a CUDA float32 tensor (this package has no CPU path)")
    cd = len(bbox)
    if not 1 <= cd <= 3 or cd > image.ndim:
        raise ValueError("crop_and_pad_nd: bbox must cover 1 to 3 trailing axes")
    lead = tuple(image.shape[:image.ndim - cd])
    target = list(lead) + [int(hi) - int(lo) for lo, hi in bbox]
    for (lo, hi), size in zip(bbox, image.shape[image.ndim - cd:]):
        if hi <= 0 or lo >= size:
            return torch.zeros(target, dtype=image.dtype, device=image.device)
    image = image.contiguous()
    dims = [1] * (3 - cd) + [int(s) for s in image.shape[image.ndim - cd:]]
    lbs = [0] * (3 - cd) + [int(lo) for lo, _ in bbox]
    patch = [1] * (3 - cd) + target[len(lead):]
    n = int(np.prod(lead)) if lead else 1
    per = dims[0] * dims[1] * dims[2]
    descs = np.zeros(n, dtype=_DESC)
    descs["src_offset"] = np.arange(n, dtype=np.int64) * per
    descs["src_dims"] = [1] + dims
    descs["lb"] = lbs
    descs["mult"] = 1.0
    dev = torch.from_numpy(descs.view(np.uint8).reshape(-1)).to(image.device)
    out = torch.empty(target, dtype=image.dtype, device=image.device)
    for b0 in range(0, n, 32768):   # grid z limit
        b1 = min(n, b0 + 32768)
        patch_gather(image, dev[b0 * _DESC.itemsize:], out.reshape(n, -1)[b0:b1], b1 - b0, 1, patch,
                     pad_value=float(pad_value))
    return out


# ------------------------------------------------------------------------------------------------------- sampler
class CustomBatchSampler(torch.utils.data.Sampler):
    """data_processing.py:601-641: `number_of_steps` batches of (position in batch, case index) per epoch; every case is
    used once before any repeats; `np.random.shuffle` as the reference."""

    def __init__(self, dataset, batch_size, number_of_steps=250, shuffle=True):
        super().__init__()
        self.batch_size = batch_size
        self.number_of_steps = number_of_steps
        self.shuffle = shuffle
        self.indices = list(range(len(dataset)))
        self.sample_order: list[int] = []

    def define_indices(self):
        if self.shuffle:
            np.random.shuffle(self.indices)
        self.sample_order = []
        available = self.indices.copy()
        while len(self.sample_order) < self.number_of_steps * self.batch_size:
            if len(available) < self.batch_size:
                available = self.indices.copy()
                if self.shuffle:
                    np.random.shuffle(available)
            self.sample_order.extend(available[:self.batch_size])
            available = available[self.batch_size:]

    def __iter__(self):
        self.define_indices()
        for step in range(self.number_of_steps):
            chunk = self.sample_order[step * self.batch_size:(step + 1) * self.batch_size]
            yield [(i, s) for i, s in enumerate(chunk)]

    def __len__(self):
        return self.number_of_steps


# ------------------------------------------------------------------------------------------------------- dataset
def _bg_contrast(rng_range):
    """batchgeneratorsv2 `BGContrast` sampler [upstream-memory]: half of the draws from [lo, 1), half from [max(lo,1), hi]."""
    lo, hi = rng_range
    if np.random.random() < 0.5 and lo < 1:
        return np.random.uniform(lo, 1)
    return np.random.uniform(max(lo, 1), hi)


def rotation_scale_matrix(angles, scales) -> np.ndarray:
    """(z, y, x) matrix M: source offset from the patch centre = M @ output offset; angles about the z, y, x axes."""
    az, ay, ax = angles
    cz, sz, cy, sy, cx, sx = math.cos(az), math.sin(az), math.cos(ay), math.sin(ay), math.cos(ax), math.sin(ax)
    rz = np.array([[1, 0, 0], [0, cz, -sz], [0, sz, cz]])
    ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    rx = np.array([[cx, -sx, 0], [sx, cx, 0], [0, 0, 1]])
    return (rz @ ry @ rx @ np.diag(np.asarray(scales, dtype=np.float64))).astype(np.float32)


class PatchSampler:
    """The host half of `MedicalDataset` (data_processing.py:274-331, 405-527): patch geometry, oversampling rule, box and
    augmentation draws -- `np.random` consumed in the reference's order. Pure host code (tested on CPU against the
    reference's goldens); `MedicalDataset` adds the resident cases and the kernels."""

    def __init__(self, batch_size, section, transformation_args, oversample_foreground_percent,
                 probabilistic_oversampling=False):
        self.batch_size = batch_size
        self.section = section
        self.transformation_args = dict(transformation_args)
        self.oversample_foreground_percent = oversample_foreground_percent
        ta = self.transformation_args
        for off in ("gaussian_noise", "gaussian_blur", "low_resolution", "dummy_2d"):
            if ta.get(off):
                raise NotImplementedError(f"transformation {off!r} is not part of the B200 data path (no planner output "
                                          "of the reference enables it, configuration.py:933-958)")
        self.patch_size = tuple(int(v) for v in ta["patch_size"])
        aug = self.configure_augmentation_params()
        self.initial_patch_size = aug["initial_patch_size"] if section == "training" else self.patch_size
        ta["rot_for_da"] = aug["rot_for_da"] if ta.get("rotation") else None
        ta["mirror_axes"] = aug["mirror_axes"] if ta.get("mirror") else None
        ta["scaling_range"] = aug["scale_range"] if ta.get("scaling") else None
        ta["brightness_range"] = aug["brightness_range"] if ta.get("brightness") else None
        ta["contrast_range"] = aug["contrast_range"] if ta.get("contrast") else None
        ta["gamma_range"] = aug["gamma_range"] if ta.get("gamma") else None
        self.is_2d = len(self.patch_size) == 2
        if self.is_2d:                        # pseudo 3-D, data_processing.py:299-302
            self.patch_size = (1, *self.patch_size)
            self.initial_patch_size = (1, *self.initial_patch_size)
        self.need_to_pad = (np.array(self.initial_patch_size) - np.array(self.patch_size)).astype(int)
        self.oversampling_method = self._oversample_last_XX_percent if not probabilistic_oversampling \
            else self._probabilistic_oversampling
        if tuple(self.initial_patch_size) != tuple(self.patch_size):
            raise NotImplementedError("initial_patch_size != patch_size (heavy augmentation) is not supported")

    # ---- host-side sampling decisions: np.random in the reference's order ---------------------------------------
    def configure_augmentation_params(self):
        """data_processing.py:405-431, the soft augmentation used for image generation."""
        dim = len(self.patch_size)
        rot_dim = 0 if dim == 3 else 2 if dim == 2 else None

        def rot(image, d):
            return np.random.uniform(-0.174533, 0.174533) if d == rot_dim else 0

        return {"rot_for_da": rot, "do_dummy_2d": False, "initial_patch_size": tuple(self.patch_size),
                "mirror_axes": (2,) if dim == 3 else (1,), "scale_range": (0.9, 1.1), "brightness_range": (0.9, 1.1),
                "contrast_range": (0.9, 1.1), "gamma_range": (0.9, 1.1)}

    def _oversample_last_XX_percent(self, sample_idx: int) -> bool:
        return sample_idx >= round(self.batch_size * (1 - self.oversample_foreground_percent))

    def _probabilistic_oversampling(self, sample_idx: int) -> bool:
        return np.random.uniform() < self.oversample_foreground_percent

    def get_bbox(self, data_shape, force_fg, class_locations, is_2d=False):
        """data_processing.py:463-527: random (or foreground-centred) box along the slice axis, jittered centre crop in
        the last two axes; bounds may lie outside the case (padding)."""
        dim = len(data_shape)
        need = self.need_to_pad.copy()
        init = self.initial_patch_size
        for d in range(dim):
            if need[d] + data_shape[d] < init[d]:
                need[d] = init[d] - data_shape[d]
        lbs = [-need[i] // 2 for i in range(dim)]
        ubs = [data_shape[i] + need[i] // 2 + need[i] % 2 - init[i] for i in range(dim)]
        box = [np.random.randint(lbs[i], ubs[i] + 1) for i in range(dim)]
        if force_fg and class_locations is not None:
            eligible = [c for c in class_locations if len(class_locations[c]) > 0]
            if eligible:
                voxels = class_locations[np.random.choice(eligible)]
                voxel = voxels[np.random.choice(len(voxels))]
                for i in range(dim):
                    if is_2d and i == 0:
                        box[0] = voxel[0]
                    elif not is_2d:
                        box[i] = max(lbs[i], min(voxel[i] - init[i] // 2, ubs[i]))
        for i in range(dim - 2, dim):
            crop, size = init[i], data_shape[i]
            center = size // 2
            if size < crop:
                box[i] = center - crop // 2
            else:
                max_offset = min(10, center - crop // 2, size - center - (crop - crop // 2))
                offset = np.random.randint(-max_offset, max_offset + 1) if max_offset > 0 else 0
                box[i] = center + offset - crop // 2
        return box, [box[i] + init[i] for i in range(dim)]

    def _draw_augmentation(self, n_channels: int) -> dict:
        """The pipeline of data_processing.py:745-858 as parameters [batchgeneratorsv2: upstream-memory]: RandomTransform
        draws first, then the transform's own parameters."""
        ta = self.transformation_args
        p = {"mat": None, "mult": None, "contrast": None, "gamma": None, "flip": [0, 0, 0]}
        if self.section != "training":
            return p
        do_rot = ta.get("rot_for_da") is not None and np.random.uniform() < 0.2
        do_scale = ta.get("scaling_range") is not None and np.random.uniform() < 0.2
        if do_rot or do_scale:
            angle = ta["rot_for_da"](None, 2 if self.is_2d else 0) if do_rot else 0.0
            scale = np.random.uniform(*ta["scaling_range"]) if do_scale else 1.0   # synchronised across axes
            p["mat"] = rotation_scale_matrix((angle, 0.0, 0.0), (1.0 if self.is_2d else scale, scale, scale))
        if ta.get("brightness_range") is not None and np.random.uniform() < 0.15:
            p["mult"] = [_bg_contrast(ta["brightness_range"]) for _ in range(n_channels)]
        if ta.get("contrast_range") is not None and np.random.uniform() < 0.15:
            p["contrast"] = [_bg_contrast(ta["contrast_range"]) for _ in range(n_channels)]
        if ta.get("gamma_range") is not None and np.random.uniform() < 0.3:
            p["gamma"] = [_bg_contrast(ta["gamma_range"]) for _ in range(n_channels)]
        if ta.get("mirror_axes"):
            for a in ta["mirror_axes"]:
                if np.random.uniform() < 0.5:
                    p["flip"][a + 1 if self.is_2d else a] = 1
        return p


class MedicalDataset(PatchSampler, torch.utils.data.Dataset):
    """Resident-in-HBM counterpart of data_processing.py:274-598 (module docstring). Extra keyword arguments:
    `device`, `out_dtype` (float32 as the reference, or bfloat16), `channels_last` (write the batch in the layout the B200
    convolutions read, (B, Z, Y, X, C) memory behind a (B, C, Z, Y, X) view), `volumes` (share an already uploaded
    `ResidentVolumes`, e.g. between the training and the validation dataset), `cases` ({name: (image, properties)} instead
    of files)."""

    def __init__(self, data_path, data_ids, batch_size, section, transformation_args, oversample_foreground_percent,
                 channel_ids=None, probabilistic_oversampling=False, *, device="cuda", out_dtype=torch.float32,
                 channels_last=False, volumes: Optional[ResidentVolumes] = None, cases: Optional[dict] = None):
        PatchSampler.__init__(self, batch_size, section, transformation_args, oversample_foreground_percent,
                              probabilistic_oversampling)
        self.data_path = data_path
        self.ids = list(data_ids)
        self.channel_ids = None if channel_ids is None else [int(c) for c in channel_ids]
        self.out_dtype, self.channels_last = out_dtype, bool(channels_last)
        _dtype_code(out_dtype)

        # the cases: uploaded once
        self.properties: list[dict] = []
        if volumes is None:
            volumes = ResidentVolumes(device)
            for name in self.ids:
                image, props = cases[name] if cases is not None else load_case(self.data_path, name)
                volumes.add(name, image)
            volumes.finalize()
        self.volumes = volumes
        self._case_of = {n: i for i, n in enumerate(volumes.names)}
        for name in self.ids:
            if cases is not None:
                self.properties.append(cases[name][1])
            else:
                with open(os.path.join(self.data_path, name + ".pkl"), "rb") as f:
                    self.properties.append(pickle.load(f))
        self._scratch = _Scratch(volumes.device)
        n_out = len(self.channel_ids) if self.channel_ids is not None else (volumes.shapes[0][0] if volumes.shapes else 1)
        if n_out > PATCH_MAX_CH:
            raise ValueError(f"at most {PATCH_MAX_CH} channels per patch")
        self.out_channels = n_out

    def __len__(self):
        return len(self.ids)

    # ---- one batch on the device --------------------------------------------------------------------------------
    def sample_batch(self, batch, return_params: bool = False):
        """`batch` = [(position in batch, index into data_ids), ...] as CustomBatchSampler yields. Returns
        {'id': [...], 'image': (B, C, Z, Y, X) (or (B, C, Y, X) for 2-D) on the device}."""
        B, Cn = len(batch), self.out_channels
        descs = np.zeros(B, dtype=_DESC)
        params, boxes = [], []
        for k, (batch_idx, sample_idx) in enumerate(batch):
            name = self.ids[sample_idx]
            case = self._case_of[name]
            shape = self.volumes.shapes[case]
            force_fg = self.oversampling_method(batch_idx)
            lbs, ubs = self.get_bbox(shape[1:], force_fg, self.properties[sample_idx].get("class_locations"),
                                     is_2d=self.patch_size[0] == 1)
            aug = self._draw_augmentation(Cn)
            d = descs[k]
            d["src_offset"] = self.volumes.offsets[case]
            d["src_dims"] = shape
            d["lb"] = [int(v) for v in lbs]
            d["flip"] = aug["flip"]
            d["affine"] = int(aug["mat"] is not None)
            d["mat"] = (aug["mat"] if aug["mat"] is not None else np.eye(3, dtype=np.float32)).reshape(-1)
            mult = np.ones(PATCH_MAX_CH, dtype=np.float32)
            if aug["mult"] is not None:
                mult[:Cn] = aug["mult"]
            d["mult"] = mult
            ch = np.zeros(PATCH_MAX_CH, dtype=np.int32)
            ch[:Cn] = self.channel_ids if self.channel_ids is not None else np.arange(Cn)
            if int(ch[:Cn].max()) >= shape[0]:
                raise IndexError(f"channel id {int(ch[:Cn].max())} out of range for case {name!r} with {shape[0]} channels")
            d["channel"] = ch
            params.append(aug)
            boxes.append((lbs, ubs))
        image = self._run(descs, params, B, Cn)
        out = {"id": [self.ids[s] for _, s in batch], "image": image}
        if return_params:
            out["params"], out["boxes"] = params, boxes
        return out

    def _run(self, descs: np.ndarray, params: list, B: int, Cn: int) -> torch.Tensor:
        dev = self.volumes.device
        P = self.patch_size
        S = P[0] * P[1] * P[2]
        with torch.cuda.device(dev):
            descs_dev = self._scratch.upload("descs", descs)
            if self.channels_last and Cn > 1:
                out = torch.empty((B, *P, Cn), dtype=self.out_dtype, device=dev)
            else:
                out = torch.empty((B, Cn, *P), dtype=self.out_dtype, device=dev)
            any_affine = any(p["mat"] is not None for p in params)
            any_contrast = any(p["contrast"] is not None for p in params)
            any_gamma = any(p["gamma"] is not None for p in params)
            cl = self.channels_last and Cn > 1
            if not (any_contrast or any_gamma):
                patch_gather(self.volumes.buffer, descs_dev, out, B, Cn, P, channels_last=cl, any_affine=any_affine,
                             clamp=(0.0, 1.0))     # data_processing.py:595
            else:
                rows = B * Cn
                stage = torch.empty((B, Cn, *P), dtype=torch.float32, device=dev)
                patch_gather(self.volumes.buffer, descs_dev, stage, B, Cn, P, any_affine=any_affine)
                ops = np.zeros((3, rows, 4), dtype=np.float32)     # contrast / gamma / retain_stats passes
                act = np.zeros((2, rows), dtype=np.int32)
                for b, p in enumerate(params):
                    for c in range(Cn):
                        if p["contrast"] is not None:
                            ops[0, b * Cn + c] = (1, p["contrast"][c], 0, 0)
                            act[0, b * Cn + c] = 1
                        if p["gamma"] is not None:
                            ops[1, b * Cn + c] = (2, p["gamma"][c], 0, 0)
                            ops[2, b * Cn + c] = (3, 0, 0, 0)
                            act[1, b * Cn + c] = 1
                ops_dev = self._scratch.upload("ops", ops).view(torch.float32)
                act_dev = self._scratch.upload("act", act).view(torch.int32)
                st = torch.zeros(3, rows, 4, dtype=torch.float32, device=dev)
                if any_contrast:
                    patch_stats(stage, rows, S, active=act_dev[:rows], out=st[0])
                    if any_gamma:
                        patch_intensity(stage, stage, ops_dev[:rows * 4], st[0], st[0], B, Cn, S)
                    else:
                        patch_intensity(stage, out, ops_dev[:rows * 4], st[0], st[0], B, Cn, S, channels_last=cl,
                                        clamp=(0.0, 1.0))
                if any_gamma:
                    patch_stats(stage, rows, S, active=act_dev[rows:2 * rows], out=st[1])
                    patch_intensity(stage, stage, ops_dev[rows * 4:rows * 8], st[1], st[1], B, Cn, S)
                    patch_stats(stage, rows, S, active=act_dev[rows:2 * rows], out=st[2])
                    patch_intensity(stage, out, ops_dev[rows * 8:rows * 12], st[1], st[2], B, Cn, S, channels_last=cl,
                                    clamp=(0.0, 1.0))
            if cl:
                out = out.permute(0, 4, 1, 2, 3)       # logical (B, C, Z, Y, X), channels-last memory
            if self.is_2d:
                out = out.squeeze(2)                    # data_processing.py:586
            return out

    def __getitem__(self, indexes):
        batch_idx, sample_idx = indexes
        one = self.sample_batch([(batch_idx, sample_idx)])
        return {"id": one["id"][0], "image": one["image"][0]}


class ResidentLoader:
    """What `DataLoader(dataset, batch_sampler=sampler)` is to the reference trainers (`for step, batch in
    enumerate(loader)`, `len(loader)`, `batch['image'].to(device)`): iterates the sampler and cuts each batch on the
    device. No workers, no pinned staging: the batch is born in HBM."""

    def __init__(self, dataset: MedicalDataset, batch_sampler: CustomBatchSampler):
        self.dataset, self.batch_sampler = dataset, batch_sampler

    def __len__(self):
        return len(self.batch_sampler)

    def __iter__(self):
        for batch in self.batch_sampler:
            yield self.dataset.sample_batch(batch)


# ------------------------------------------------------------------------------------------- splits (host, :34-147)
def generate_crossval_split(train_identifiers, seed=12345, n_splits=5):
    from sklearn.model_selection import KFold
    ids = np.array(train_identifiers)
    return [{"train": list(ids[tr]), "val": list(ids[te])}
            for tr, te in KFold(n_splits=n_splits, shuffle=True, random_state=seed).split(train_identifiers)]


def create_split_files(dataset_id, splitting, model_type, seed=12345):
    """data_processing.py:47-101: `splits_train_val_test.json` (70/10/20) or `splits_final.json` (5-fold) next to
    `imagesTr` under $medimgen_preprocessed/Task<id>*; an existing file is reused."""
    from sklearn.model_selection import train_test_split
    root = glob.glob(os.getenv("medimgen_preprocessed") + f"/Task{dataset_id}*/")[0]
    images = os.path.join(root, "imagesTr")
    split_path = os.path.join(root, "splits_train_val_test.json" if splitting == "train-val-test" else "splits_final.json")
    if os.path.exists(split_path):
        return split_path
    names = []
    for ext in (".zarr", ".npz", ".b2nd"):
        names = [os.path.basename(p)[:-len(ext)] for p in glob.glob(os.path.join(images, "*" + ext))
                 if not (ext == ".b2nd" and "_seg" in p)]
        if names:
            break
    if splitting == "train-val-test":
        train_val, test = train_test_split(names, test_size=0.2, random_state=seed)
        train, val = train_test_split(train_val, test_size=0.125, random_state=seed)
        split = {"train": train, "val": val, "test": test}
    elif splitting == "5-fold":
        split = generate_crossval_split(names, seed=seed, n_splits=5)
    else:
        raise ValueError("Invalid splitting option. Choose 'train-val-test' or '5-fold'.")
    with open(split_path, "w") as f:
        json.dump(split, f, indent=4)
    return split_path


def get_data_ids(split_file_path, fold=None):
    with open(split_file_path) as f:
        split = json.load(f)
    part = split[int(fold)] if fold is not None else split
    return {"train": part["train"], "val": part["val"]}


def get_data_loaders(config, dataset_id, splitting, batch_size, model_type, transformations, fold=None, *,
                     device="cuda", out_dtype=torch.float32, channels_last=False):
    """data_processing.py:118-147 with the datasets resident in HBM: (train_loader, val_loader), 250 / 50 steps."""
    split_path = create_split_files(dataset_id, splitting, model_type, seed=12345)
    ids = get_data_ids(split_path, fold)
    root = glob.glob(os.getenv("medimgen_preprocessed") + f"/Task{dataset_id}*/")[0]
    images = os.path.join(root, "imagesTr")
    kw = dict(oversample_foreground_percent=config["oversample_ratio"], channel_ids=config["input_channels"],
              device=device, out_dtype=out_dtype, channels_last=channels_last)
    train_ds = MedicalDataset(images, ids["train"], batch_size, "training", transformations, **kw)
    val_ds = MedicalDataset(images, ids["val"], batch_size, "validation", transformations, **kw)
    return (ResidentLoader(train_ds, CustomBatchSampler(train_ds, batch_size, number_of_steps=250, shuffle=True)),
            ResidentLoader(val_ds, CustomBatchSampler(val_ds, batch_size, number_of_steps=50, shuffle=False)))
