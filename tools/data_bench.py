"""Data path on one B200: batches/s of the resident-in-HBM loader through its public API (host RNG +
descriptor upload + kernels, result on the device) at the BASELINE config-5 patch, the kernels' GB/s against the measured
copy bandwidth, and the reference's per-sample path (numpy crop + pad + clamp restated in oracle/data_oracle.py, plus the
host->device copy the trainers do) on the host cores beside it.

    python tools/data_bench.py [--probe gather|stats|intensity]     (--probe: few launches of one kernel, for ncu)"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from medical_image_generation_b200 import data as mdata  # noqa: E402

OFF = {"scaling": False, "rotation": False, "gaussian_noise": False, "gaussian_blur": False, "low_resolution": False,
       "brightness": False, "contrast": False, "gamma": False, "mirror": False, "dummy_2d": False}
ON = dict(OFF, scaling=True, rotation=True, brightness=True, contrast=True, gamma=True, mirror=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--probe", default=None)
    ap.add_argument("--cases", type=int, default=24)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    pk = bench.peaks()
    patch, B = (160, 160, 128), 2
    rs = np.random.RandomState(0)
    shape = (2, 176, 176, 150)
    cases = {}
    g = torch.Generator(device=dev).manual_seed(0)
    for i in range(args.cases):
        locs = {1: [tuple(int(rs.randint(s)) for s in shape[1:]) for _ in range(50)]}
        cases[f"c{i}"] = (torch.rand(shape, generator=g, device=dev), {"class_locations": locs})
    out = {}
    for label, tf in (("plain", OFF), ("augmented", ON)):
        ds = mdata.MedicalDataset("", sorted(cases), B, "training", dict(tf, patch_size=list(patch)), 0.33, cases=cases)
        loader = mdata.ResidentLoader(ds, mdata.CustomBatchSampler(ds, B, number_of_steps=60, shuffle=True))
        if args.probe:
            it = iter(loader)
            for _ in range(4):
                next(it)
            torch.cuda.synchronize()
            if label == "augmented":
                return
            continue
        np.random.seed(0)
        for _ in loader:      # warm-up epoch (allocator, pinned tables)
            pass
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n = 0
        for batch in loader:
            n += batch["image"].shape[0]
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        nbytes = 2 * n * 2 * patch[0] * patch[1] * patch[2] * 4
        out[label] = {"patches_per_s": n / dt, "ms_per_batch": 1e3 * dt / (n / B), "algorithmic_GBps_plain_crop": nbytes / dt / 1e9}
    # the reference's path per sample on the host: numpy crop + pad + clamp (transforms off), collate, pin, H2D
    from oracle import data_oracle as D
    host_cases = [(c[0].cpu().numpy(), c[1]["class_locations"]) for c in list(cases.values())[:4]]
    np.random.seed(0)
    t0 = time.perf_counter()
    n = 0
    while time.perf_counter() - t0 < 10.0:
        imgs = [D.getitem_untransformed(*host_cases[(n + k) % 4], k, B, patch, 0.33) for k in range(B)]
        batch = torch.from_numpy(np.stack(imgs)).pin_memory().to(dev, non_blocking=True)
        torch.cuda.synchronize()
        n += B
    dt = time.perf_counter() - t0
    out["cpu_reference_path"] = {"patches_per_s": n / dt, "cores": 1, "kind": "port",
                                 "sample": f"{n} patches of 2x160x160x128 in {dt:.1f} s: numpy crop + pad + clamp, stack, pin, H2D"}
    hb = bench.hbm_block(dev, pk)
    out["kernels"] = {k: v for k, v in hb["kernels"].items() if k.startswith("patch_")}
    out["hbm_peak_gbs"] = pk["hbm"]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
