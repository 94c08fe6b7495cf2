"""GPU parity of the data path (csrc/data_path.cu through the C ABI, medical_image_generation_b200/data.py) against
oracle/data_oracle.py and the reference-generated goldens (tests/golden/data_path.pt).

Bar: crop / pad / channel selection / mirror / clamp are byte movement -> BIT-EXACT (sha256 of the fp32 bytes the
unmodified reference produced). Intensity transforms are fp32 arithmetic -> 2e-5 absolute on data in [0, 1]; resampled
patches 2e-4 (torch's own grid_sample / mean / std / pow are the oracle's arithmetic)."""
import hashlib

import numpy as np
import pytest
import torch

from medical_image_generation_b200 import data as pkg
from oracle import data_oracle as D
from oracle.gen_golden_data import DATASETS, make_case

pytestmark = pytest.mark.gpu

_OFF = {"scaling": False, "rotation": False, "gaussian_noise": False, "gaussian_blur": False, "low_resolution": False,
        "brightness": False, "contrast": False, "gamma": False, "mirror": False, "dummy_2d": False}
_ON = dict(_OFF, scaling=True, rotation=True, brightness=True, contrast=True, gamma=True, mirror=True)


def _cases(rec):
    return {f"case{i}": (img, {"class_locations": locs})
            for i, (img, locs) in enumerate(make_case(s, shp) for s, shp in zip(rec["seeds"], rec["shapes"]))}


def test_crop_and_pad_kernel_matches_reference_golden(golden):
    for case in golden("data_path")["crop"]:
        got = pkg.crop_and_pad_nd(case["image"].cuda(), case["bbox"], case["pad"])
        assert got.shape == case["out"].shape
        assert torch.equal(got.cpu(), case["out"]), case["bbox"]


def test_crop_and_pad_kernel_random_boxes_vs_oracle():
    rs = np.random.RandomState(5)
    for _ in range(40):
        nd = int(rs.randint(1, 5))
        shape = tuple(int(v) for v in rs.randint(1, 20, size=nd))
        cd = int(rs.randint(1, min(nd, 3) + 1))
        img = rs.rand(*shape).astype(np.float32)
        bbox = [[int(lo), int(lo + rs.randint(1, 24))] for lo in rs.randint(-10, 20, size=cd)]
        pad = float(rs.choice([0.0, -1.0, 0.5]))
        want = D.crop_and_pad_nd(img, bbox, pad)
        got = pkg.crop_and_pad_nd(torch.from_numpy(img).cuda(), bbox, pad)
        assert np.array_equal(got.cpu().numpy(), want), (shape, bbox, pad)


@pytest.mark.parametrize("name", sorted(DATASETS))
@pytest.mark.parametrize("section", ["training", "validation"])
def test_dataset_matches_reference_golden_bit_exact(golden, name, section):
    """Seeded sampler + boxes + crop + pad + channel selection + clamp = the unmodified MedicalDataset.__getitem__."""
    rec = golden("data_path")["datasets"][name]
    want = rec[section]
    ds = pkg.MedicalDataset("", [f"case{i}" for i in range(len(rec["shapes"]))], rec["batch_size"], section,
                            dict(_OFF, patch_size=list(rec["patch"])), rec["oversample"], channel_ids=rec["channel_ids"],
                            cases=_cases(rec))
    loader = pkg.ResidentLoader(ds, pkg.CustomBatchSampler(ds, rec["batch_size"], number_of_steps=3,
                                                          shuffle=section == "training"))
    assert len(loader) == 3
    np.random.seed(2024)
    k = 0
    for step, batch in enumerate(loader):
        img = batch["image"]
        assert img.is_cuda and img.dtype == torch.float32 and tuple(img.shape[1:]) == tuple(want["shape"])
        assert batch["id"] == [f"case{s}" for _, s in want["batches"][step]]
        host = img.cpu().numpy()
        for b in range(host.shape[0]):
            assert hashlib.sha256(np.ascontiguousarray(host[b]).tobytes()).hexdigest() == want["sha256"][k], (step, b)
            if k < rec["batch_size"]:
                assert np.array_equal(host[b], want["images"][k].numpy())
            k += 1
    assert k == len(want["sha256"])


def test_getitem_single_sample_like_reference(golden):
    rec = golden("data_path")["datasets"]["3d_fits"]
    ds = pkg.MedicalDataset("", ["case0", "case1", "case2"], rec["batch_size"], "validation",
                            dict(_OFF, patch_size=list(rec["patch"])), rec["oversample"], cases=_cases(rec))
    np.random.seed(2024)
    pos, idx = rec["validation"]["batches"][0][0]
    item = ds[(pos, idx)]
    assert item["id"] == f"case{idx}"
    assert np.array_equal(item["image"].cpu().numpy(), rec["validation"]["images"][0].numpy())


@pytest.mark.parametrize("channels_last", [False, True])
def test_bf16_and_channels_last_outputs_are_the_rounded_fp32_batch(golden, channels_last):
    rec = golden("data_path")["datasets"]["3d_fits"]
    kw = dict(cases=_cases(rec))
    args = ("", ["case0", "case1", "case2"], rec["batch_size"], "training", dict(_OFF, patch_size=list(rec["patch"])),
            rec["oversample"])
    batch = [(i, i % 3) for i in range(4)]
    np.random.seed(1)
    ref32 = pkg.MedicalDataset(*args, **kw).sample_batch(batch)["image"]
    np.random.seed(1)
    got = pkg.MedicalDataset(*args, out_dtype=torch.bfloat16, channels_last=channels_last, **kw).sample_batch(batch)["image"]
    assert got.dtype == torch.bfloat16 and got.shape == ref32.shape
    if channels_last:
        assert got.is_contiguous(memory_format=torch.channels_last_3d)
    assert torch.equal(got.float(), ref32.to(torch.bfloat16).float())


def _oracle_pipeline(image, box, aug, channel_ids):
    lbs, ubs = box
    x = D.crop_and_pad_nd(np.asarray(image), [[int(a), int(b)] for a, b in zip(lbs, ubs)], 0)
    if channel_ids is not None:
        x = x[channel_ids]
    x = x.astype(np.float32)
    if aug["mat"] is not None:
        x = D.affine_resample(x, aug["mat"])
    if aug["mult"] is not None:
        x = D.brightness(x, aug["mult"])
    if aug["contrast"] is not None:
        x = D.contrast(x, aug["contrast"])
    if aug["gamma"] is not None:
        x = D.gamma(x, aug["gamma"], invert=False, retain_stats=True)
    x = D.mirror(x, [a for a in range(3) if aug["flip"][a]])
    return np.clip(x, 0.0, 1.0)


@pytest.mark.parametrize("name", ["3d_fits", "3d_needs_padding", "3d_channel_select", "2d"])
def test_augmented_batches_match_oracle_pipeline(golden, name):
    """Every transform the planner switches on, with the parameters the dataset drew, replayed through the oracle."""
    rec = golden("data_path")["datasets"][name]
    cases = _cases(rec)
    ids = sorted(cases)
    ds = pkg.MedicalDataset("", ids, 8, "training", dict(_ON, patch_size=list(rec["patch"])), 0.33,
                            channel_ids=rec["channel_ids"], cases=cases)
    np.random.seed(77)
    seen = {"mat": 0, "mult": 0, "contrast": 0, "gamma": 0, "flip": 0, "contrast+gamma": 0}
    for step in range(10):
        batch = [(i, int(np.random.randint(len(ids)))) for i in range(8)]
        out = ds.sample_batch(batch, return_params=True)
        host = out["image"].cpu().numpy()
        for b, ((_, idx), aug, box) in enumerate(zip(batch, out["params"], out["boxes"])):
            want = _oracle_pipeline(cases[ids[idx]][0], box, aug, rec["channel_ids"])
            got = host[b][:, None] if ds.is_2d else host[b]
            assert got.shape == want.shape
            err = float(np.abs(got - want).max())
            # resampled patches: grid_sample rebuilds the voxel coordinate from a normalised fp32 grid (~1e-5 voxel of
            # rounding at these sizes, times a slope of up to 1.4 per voxel in this random data)
            tol = 2e-4 if aug["mat"] is not None else 2e-5
            assert err < tol, (name, step, b, err, {k: v is not None for k, v in aug.items()})
            for key in ("mat", "mult", "contrast", "gamma"):
                seen[key] += aug[key] is not None
            seen["flip"] += any(aug["flip"])
            seen["contrast+gamma"] += aug["contrast"] is not None and aug["gamma"] is not None
    assert all(v > 0 for k, v in seen.items() if k != "contrast+gamma"), seen


def test_intensity_kernels_vs_oracle_including_inverted_gamma():
    rs = np.random.RandomState(2)
    B, Cn, P = 3, 2, (6, 20, 33)
    S = int(np.prod(P))
    x = (rs.rand(B, Cn, *P) * 1.3 - 0.1).astype(np.float32)
    xd = torch.from_numpy(x).cuda()
    rows = B * Cn
    st = pkg.patch_stats(xd, rows, S)
    t = torch.from_numpy(x).reshape(rows, S)
    want = torch.stack([t.mean(1), t.std(1), t.min(1).values, t.max(1).values], 1)
    assert torch.allclose(st.cpu(), want, rtol=1e-5, atol=1e-7)
    # skipped rows are left untouched
    act = torch.tensor([1, 0, 1, 0, 1, 0], dtype=torch.int32, device="cuda")
    st2 = pkg.patch_stats(xd, rows, S, active=act, out=torch.full((rows, 4), -7.0, device="cuda"))
    assert torch.equal(st2[1::2].cpu(), torch.full((3, 4), -7.0)) and torch.equal(st2[0::2], st[0::2])
    gam = [0.8, 1.2, None, 0.95, 1.05, None]
    for invert in (False, True):
        op = np.zeros((rows, 4), dtype=np.float32)
        for r, g in enumerate(gam):
            if g is not None:
                op[r] = (2, g, float(invert), 0)
        opd = torch.from_numpy(op).cuda().reshape(-1)
        y = torch.empty_like(xd)
        pkg.patch_intensity(xd, y, opd, st, st, B, Cn, S)
        st_after = pkg.patch_stats(y, rows, S)
        op[:, 0] = np.where(op[:, 0] == 2, 3, 0)
        z = torch.empty_like(xd)
        pkg.patch_intensity(y, z, torch.from_numpy(op).cuda().reshape(-1), st, st_after, B, Cn, S, clamp=(0.0, 1.0))
        for b in range(B):
            want_b = np.clip(D.gamma(x[b], gam[b * Cn:(b + 1) * Cn], invert=invert, retain_stats=True), 0, 1)
            assert float(np.abs(z[b].cpu().numpy() - want_b).max()) < 2e-5, (invert, b)


def test_full_size_properties_config5_patch():
    """BASELINE config 5 patch (2 x 160 x 160 x 128, batch 2): size-independent properties, bit-exact."""
    g = torch.Generator(device="cuda").manual_seed(0)
    vols = pkg.ResidentVolumes("cuda")
    shapes = [(2, 170, 180, 150), (2, 150, 168, 128)]
    host = {}
    for i, shp in enumerate(shapes):
        host[i] = torch.rand(shp, generator=g, device="cuda") * 1.4 - 0.2
        vols.add(f"c{i}", host[i])
    vols.finalize()
    P = (160, 160, 128)
    descs = np.zeros(2, dtype=pkg._DESC)
    lbs = [(3, 7, 11), (-6, 4, -5)]                   # the second box sticks out of the case on two axes
    for b in range(2):
        descs[b]["src_offset"] = vols.offsets[b]
        descs[b]["src_dims"] = shapes[b]
        descs[b]["lb"] = lbs[b]
        descs[b]["mult"] = 1.0
        descs[b]["mat"] = np.eye(3, dtype=np.float32).reshape(-1)
        descs[b]["channel"][:2] = (0, 1)
    up = lambda d: torch.from_numpy(d.view(np.uint8).reshape(-1).copy()).cuda()   # noqa: E731
    out = torch.empty(2, 2, *P, device="cuda")
    pkg.patch_gather(vols.buffer, up(descs), out, 2, 2, P)
    # (1) equals torch slicing + zero padding
    assert torch.equal(out[0], host[0][:, 3:163, 7:167, 11:139])
    want1 = torch.zeros(2, *P, device="cuda")
    want1[:, 6:156, :160, 5:128] = host[1][:, 0:150, 4:164, 0:123]
    assert torch.equal(out[1], want1)
    # (2) mirror == torch.flip, and mirroring twice through two boxes is the identity
    descs["flip"] = (1, 0, 1)
    flipped = torch.empty_like(out)
    pkg.patch_gather(vols.buffer, up(descs), flipped, 2, 2, P)
    assert torch.equal(flipped, out.flip(2, 4))
    # (3) clamp is idempotent and equals torch.clamp; the affine instantiation with an identity matrix changes nothing
    descs["flip"] = 0
    clamped = torch.empty_like(out)
    pkg.patch_gather(vols.buffer, up(descs), clamped, 2, 2, P, clamp=(0.0, 1.0))
    assert torch.equal(clamped, out.clamp(0, 1))
    descs["affine"] = 1
    ident = torch.empty_like(out)
    pkg.patch_gather(vols.buffer, up(descs), ident, 2, 2, P, any_affine=True)
    assert torch.equal(ident, out)
    # (4) statistics over 3.3 M voxels per row against torch in fp64
    st = pkg.patch_stats(out, 4, out[0, 0].numel())
    flat = out.reshape(4, -1).double()
    want = torch.stack([flat.mean(1), flat.std(1), flat.min(1).values, flat.max(1).values], 1).float()
    assert torch.allclose(st, want, rtol=5e-6, atol=1e-7)
    # (5) channels-last bf16 output holds the same values
    cl = torch.empty(2, *P, 2, device="cuda", dtype=torch.bfloat16)
    descs["affine"] = 0
    descs["lb"][1] = lbs[1]
    pkg.patch_gather(vols.buffer, up(descs), cl, 2, 2, P, channels_last=True)
    assert torch.equal(cl.permute(0, 4, 1, 2, 3).float(), out.to(torch.bfloat16).float())


def test_kernel_argument_errors_are_reported():
    x = torch.zeros(8, device="cuda")
    d = torch.zeros(152, dtype=torch.uint8, device="cuda")
    with pytest.raises(RuntimeError, match="channels"):
        pkg.patch_gather(x, d, x, 1, 9, (1, 1, 1))
    with pytest.raises(RuntimeError, match="in-place"):
        pkg.patch_intensity(x, x, x, x, x, 1, 2, 4, channels_last=True)


def test_get_data_loaders_from_files_feeds_the_autoencoder(tmp_path, monkeypatch):
    """Files -> HBM -> batches -> B200 AutoencoderKL: zarr v2 cases written by hand (zlib chunks of (1, 1, Y, X), the
    chunking of configuration.py:1404-1410), split file created like data_processing.py:47-101, 250 / 50 step loaders."""
    import json
    import pickle
    from test_data_oracle import _write_zarr
    import medical_image_generation_b200 as mig
    root = tmp_path / "Task001_Synthetic"
    images = root / "imagesTr"
    images.mkdir(parents=True)
    rs = np.random.RandomState(0)
    truth = {}
    for i in range(10):
        arr = rs.rand(1, 20 + i, 40, 36).astype(np.float32)
        truth[f"p{i:02d}"] = arr
        _write_zarr(str(images / f"p{i:02d}.zarr" / "image"), arr, (1, 1, 40, 36), {"id": "zlib", "level": 1})
        # _write_zarr leaves chunk #1 out on purpose (fill value 0)
        truth[f"p{i:02d}"] = arr.copy()
        truth[f"p{i:02d}"][0, 1] = 0
        with open(images / f"p{i:02d}.pkl", "wb") as f:
            pickle.dump({"class_locations": {1: [(5, 20, 18)]}}, f)
    monkeypatch.setenv("medimgen_preprocessed", str(tmp_path))
    config = {"oversample_ratio": 0.33, "input_channels": [0], "num_workers": 0}
    tf = dict(_ON, patch_size=[16, 32, 32])
    np.random.seed(0)
    train, val = pkg.get_data_loaders(config, "001", "train-val-test", 2, "3d", tf)
    split = json.loads((root / "splits_train_val_test.json").read_text())
    assert len(split["train"]) == 7 and len(split["val"]) == 1 and len(split["test"]) == 2
    assert len(train) == 250 and len(val) == 50
    assert sorted(train.dataset.ids) == sorted(split["train"])
    # the resident buffer holds the files' voxels
    vols = train.dataset.volumes
    for name, off, shp in zip(vols.names, vols.offsets, vols.shapes):
        n = int(np.prod(shp))
        assert np.array_equal(vols.buffer[off:off + n].cpu().numpy().reshape(shp), truth[name])
    ae = mig.AutoencoderKL(spatial_dims=3, in_channels=1, out_channels=1, num_res_blocks=1, num_channels=(16, 32),
                           attention_levels=(False, False), latent_channels=3, norm_num_groups=8,
                           with_encoder_nonlocal_attn=False, with_decoder_nonlocal_attn=False,
                           downsample_parameters=[[[1, 1, 1], [3, 3, 3], [1, 1, 1]], [[2, 2, 2], [3, 3, 3], [1, 1, 1]]],
                           upsample_parameters=[[[2, 2, 2], [3, 3, 3], [1, 1, 1]]]).cuda()
    for step, batch in enumerate(train):
        img = batch["image"]
        assert img.shape == (2, 1, 16, 32, 32) and img.is_cuda and len(batch["id"]) == 2
        assert float(img.min()) >= 0.0 and float(img.max()) <= 1.0
        recon, mu, sigma = ae(img)
        assert recon.shape == img.shape and torch.isfinite(recon.float()).all()
        if step == 3:
            break
    vb = next(iter(val))
    assert vb["image"].shape == (2, 1, 16, 32, 32)
