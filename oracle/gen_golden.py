"""Generate tests/golden/*.pt by running the UNMODIFIED reference (under the MONAI shim) on CPU fp32.

Build-container only (needs /root/reference).  python -m oracle.gen_golden
Each golden holds: case cfg, param seed, (key -> shape) of the reference state_dict, inputs,
reference outputs, a few intermediate block outputs, input gradient, per-parameter gradient sketches.
"""
from __future__ import annotations

import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import reference_loader as ref  # noqa: E402
from oracle.golden_util import CASES, golden_inputs, golden_params, sketch  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
SEED = 20261018


def run_case(name: str, case: dict) -> dict:
    torch.manual_seed(0)
    if case["kind"] == "unet":
        model = ref.unet_module().DiffusionModelUNet(**case["cfg"])
    else:
        model = ref.ae_module().AutoencoderKL(**case["cfg"])
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    params = golden_params(shapes, SEED)
    model.load_state_dict(params)
    model.train()
    inp = golden_inputs(case, SEED)
    g = {"name": name, "kind": case["kind"], "cfg": case["cfg"], "batch": case["batch"], "seed": SEED,
         "shapes": shapes, "torch": torch.__version__}
    x = inp["x"].clone().requires_grad_(True)
    if case["kind"] == "unet":
        taps = {}
        hooks = []
        watch = {"conv_in": model.conv_in, "down0_res0": model.down_blocks[0].resnets[0],
                 "mid": model.middle_block, "up0": model.up_blocks[0], "mid_attn": model.middle_block.attention}
        for k, m in watch.items():
            hooks.append(m.register_forward_hook(lambda _m, _i, o, k=k: taps.__setitem__(k, o.detach().clone())))
        kw = {}
        if "context" in inp:
            kw = dict(context=inp["context"], class_labels=inp["class_labels"])
        y = model(x, inp["timesteps"], **kw)
        (y * inp["probe"]).sum().backward()
        for h in hooks:
            h.remove()
        g["inputs"] = {k: v for k, v in inp.items()}
        g["out"] = y.detach().clone()
        g["taps"] = taps
    else:
        z_mu, z_sigma = model.encode(x)
        gen = torch.Generator().manual_seed(SEED + 1)
        eps = torch.randn(z_sigma.shape, generator=gen)
        z = z_mu + eps * z_sigma
        recon = model.decode(z)
        kl = 0.5 * torch.sum(z_mu.pow(2) + z_sigma.pow(2) - torch.log(z_sigma.pow(2)) - 1,
                             dim=list(range(1, z_sigma.ndim)))
        kl = torch.sum(kl) / kl.shape[0]          # train_autoencoder.py:67-72
        loss = torch.nn.functional.l1_loss(recon.float(), inp["x"].float()) + 1e-7 * kl  # :412-414, cfg :1020
        loss.backward()
        inp["eps"] = eps
        g["inputs"] = inp
        g["out"] = recon.detach().clone()
        g["z_mu"], g["z_sigma"] = z_mu.detach().clone(), z_sigma.detach().clone()
        g["loss"] = loss.detach().clone()
        g["kl"] = kl.detach().clone()
    g["grad_x"] = x.grad.detach().clone()
    g["grad_sketch"] = {k: sketch(p.grad) for k, p in model.named_parameters() if p.grad is not None}
    g["no_grad_params"] = sorted(k for k, p in model.named_parameters() if p.grad is None)
    return g


def main():
    os.makedirs(OUT, exist_ok=True)
    only = [a for a in sys.argv[1:] if not a.startswith("-")]   # python -m oracle.gen_golden [case ...]
    for name, case in CASES.items():
        if only and name not in only:
            continue
        g = run_case(name, case)
        path = os.path.join(OUT, name + ".pt")
        torch.save(g, path)
        print(f"{name}: out {tuple(g['out'].shape)} |out|max {float(g['out'].abs().max()):.4f} "
              f"params {sum(torch.Size(s).numel() for s in g['shapes'].values())} -> {os.path.getsize(path)/1e3:.0f} kB")
    if only and "planner" not in only:
        return
    # planner goldens: the four pure functions of configuration.py:751-902
    pf = ref.planner_functions()
    sizes = [[96, 96, 96], [24, 24, 24], [128, 128, 64], [160, 160, 128], [64, 64], [32, 32, 16], [40, 40, 32],
             [48, 96, 96], [256, 256, 32]]
    plan = {"sizes": sizes, "params": {}, "out": {}}
    for s in sizes:
        for n in (2, 3, 4):
            p = pf["compute_downsample_parameters"](list(s), n)
            plan["params"][(tuple(s), n)] = p
            plan["out"][(tuple(s), n)] = pf["compute_output_size"](list(s), p)
    torch.save(plan, os.path.join(OUT, "planner.pt"))
    print("planner:", len(plan["params"]), "entries")


if __name__ == "__main__":
    main()
