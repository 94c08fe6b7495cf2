def ensure_tuple_rep(tup, dim):
    """(x,)*dim for scalars; pass-through for length-`dim` sequences."""
    if isinstance(tup, (list, tuple)):
        if len(tup) == dim:
            return tuple(tup)
        raise ValueError(f"Sequence must have length {dim}, got {len(tup)}.")
    return (tup,) * dim


def set_determinism(seed=None, **_kwargs):
    """monai.utils.set_determinism (imported by train_ldm.py:27): seed python / numpy / torch."""
    import random

    import numpy as np
    import torch
    if seed is not None:
        random.seed(seed)
        np.random.seed(seed % (2 ** 32))
        torch.manual_seed(seed)
