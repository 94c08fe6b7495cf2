"""Restatement of `monai.networks.blocks.Convolution` (conv_only=True form) and `MLPBlock` (GEGLU form)."""
from __future__ import annotations

import torch
from torch import nn


def _tup(v, n):
    if isinstance(v, (list, tuple)):
        assert len(v) == n, (v, n)
        return tuple(int(i) for i in v)
    return (int(v),) * n


class Convolution(nn.Sequential):
    """`nn.Sequential` whose only child is named ``conv`` (hence the ``.conv.`` level in state_dict keys).

    padding=None means same-padding (k-1)//2; transposed convs use output_padding = stride-1.
    Only conv_only=True is used by the reference (unet:510-518, ae:66-129).
    """

    def __init__(self, spatial_dims, in_channels, out_channels, strides=1, kernel_size=3, padding=None,
                 conv_only=True, is_transposed=False, bias=True, **unused):
        super().__init__()
        assert conv_only, "the reference only builds conv_only=True blocks"
        k = _tup(kernel_size, spatial_dims)
        s = _tup(strides, spatial_dims)
        p = tuple((ki - 1) // 2 for ki in k) if padding is None else _tup(padding, spatial_dims)
        if is_transposed:
            cls = {2: nn.ConvTranspose2d, 3: nn.ConvTranspose3d}[spatial_dims]
            conv = cls(in_channels, out_channels, k, s, p, output_padding=tuple(si - 1 for si in s), bias=bias)
        else:
            cls = {1: nn.Conv1d, 2: nn.Conv2d, 3: nn.Conv3d}[spatial_dims]
            conv = cls(in_channels, out_channels, k, s, p, bias=bias)
        self.add_module("conv", conv)


class MLPBlock(nn.Module):
    """GEGLU feed-forward: linear1 -> (x, gate) -> x*gelu(gate) -> drop -> linear2 -> drop."""

    def __init__(self, hidden_size, mlp_dim, dropout_rate=0.0, act="GEGLU", **unused):
        super().__init__()
        assert act == "GEGLU", "the reference only uses GEGLU (unet:213)"
        self.linear1 = nn.Linear(hidden_size, mlp_dim * 2)
        self.linear2 = nn.Linear(mlp_dim, hidden_size)
        self.drop1 = nn.Dropout(dropout_rate)
        self.drop2 = nn.Dropout(dropout_rate)

    def forward(self, x):
        x, gate = self.linear1(x).chunk(2, dim=-1)
        x = self.drop1(x * torch.nn.functional.gelu(gate))
        return self.drop2(self.linear2(x))
