"""Per-parameter gradient error of the CUDA U-Net/AE vs the CPU oracle on a golden case (diagnostic)."""
import sys

import torch

sys.path.insert(0, ".")
import medical_image_generation_b200 as mig  # noqa: E402
from oracle import torch_oracle as O  # noqa: E402
from oracle.golden_util import golden_params  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "unet3d_small"
dtype = torch.bfloat16 if (len(sys.argv) < 3 or sys.argv[2] == "bf16") else torch.float32
g = torch.load(f"tests/golden/{name}.pt", weights_only=False)
params = golden_params(g["shapes"], g["seed"])
ref = {k: v.clone().requires_grad_(True) for k, v in params.items()}
inp = g["inputs"]
y_ref = O.unet_forward(ref, g["cfg"], inp["x"], inp["timesteps"], context=inp.get("context"),
                       class_labels=inp.get("class_labels"))
(y_ref * inp["probe"]).sum().backward()
m = mig.DiffusionModelUNet(**g["cfg"], compute_dtype=dtype)
m.load_state_dict(params)
m = m.cuda().train()
kw = {}
if "context" in inp:
    kw = dict(context=inp["context"].cuda(), class_labels=inp["class_labels"].cuda())
y = m(inp["x"].cuda(), inp["timesteps"].cuda(), **kw)
(y * inp["probe"].cuda()).sum().backward()
rows = []
for k, p in m.named_parameters():
    if p.grad is None:
        continue
    a, b = p.grad.detach().double().cpu(), ref[k].grad.double()
    rows.append(((a - b).norm().item() / max(b.norm().item(), 1e-30), b.norm().item(), k))
rows.sort(reverse=True)
print("out rel err", float((y.cpu().double() - y_ref.double()).norm() / y_ref.double().norm()))
for r in rows[:25]:
    print(f"{r[0]:.3e}  |ref|={r[1]:.3e}  {r[2]}")
print("median", sorted(r[0] for r in rows)[len(rows) // 2])
