"""`Pool[Pool.AVG, n]` -> nn.AvgPool{n}d (the only use is unet:522)."""
from torch import nn


class _PoolFactory:
    AVG = "avg"
    MAX = "max"

    def __getitem__(self, key):
        kind, dims = key
        table = {("avg", 1): nn.AvgPool1d, ("avg", 2): nn.AvgPool2d, ("avg", 3): nn.AvgPool3d,
                 ("max", 1): nn.MaxPool1d, ("max", 2): nn.MaxPool2d, ("max", 3): nn.MaxPool3d}
        return table[(kind, dims)]


Pool = _PoolFactory()
