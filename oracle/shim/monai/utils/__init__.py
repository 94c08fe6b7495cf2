def ensure_tuple_rep(tup, dim):
    """(x,)*dim for scalars; pass-through for length-`dim` sequences."""
    if isinstance(tup, (list, tuple)):
        if len(tup) == dim:
            return tuple(tup)
        raise ValueError(f"Sequence must have length {dim}, got {len(tup)}.")
    return (tup,) * dim
