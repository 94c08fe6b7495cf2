"""TEST SCAFFOLDING: stand-in for torchinfo.summary (absent from this image). Like the real one it performs a dry-run
forward pass of `model` on random tensors of `input_size` (train_ldm.py:529-530) -- the part that touches the hot path."""
import torch


def summary(model, input_size=None, batch_dim=None, depth=3, **_k):
    sizes = input_size if isinstance(input_size, list) else [input_size]
    dev = next(model.parameters()).device
    was_training = model.training
    model.eval()
    with torch.no_grad():
        out = model(*[torch.rand(*s, device=dev) for s in sizes])
    model.train(was_training)
    n = sum(p.numel() for p in model.parameters())
    print(f"[torchinfo stub] {type(model).__name__}: {n} parameters, dry-run forward ok")
    return out
