"""Parameter-holding leaf modules shared by the U-Net and the autoencoder.

They reproduce the reference's state_dict key layout:
  monai `Convolution(conv_only=True)`  ->  `<name>.conv.{weight,bias}`      (ConvBlock -> ConvNd)
  `nn.GroupNorm`                       ->  `<name>.{weight,bias}`           (GroupNorm)
  `nn.Linear`                          ->  `<name>.{weight,bias}`           (Linear)
but every forward goes through medical_image_generation_b200.ops (hand-written sm_100a kernels).
Conv filters are kept in channels-last memory ([Cout][taps][Cin]) which is what the implicit-GEMM
kernels read; the logical shape stays (Cout, Cin, *k) so checkpoints load and save unchanged.
"""
from __future__ import annotations

import math
from typing import Sequence

import torch
from torch import nn

from . import ops


def _tup(v, n: int) -> tuple:
    if isinstance(v, (list, tuple)):
        if len(v) != n:
            raise ValueError(f"expected {n} values, got {v}")
        return tuple(int(i) for i in v)
    return (int(v),) * n


def _cl_format(spatial_dims: int):
    return torch.channels_last_3d if spatial_dims == 3 else torch.channels_last


class ConvNd(nn.Module):
    """nn.Conv{2,3}d replacement: same parameters / default initialisation, kernels from ops.conv_nd."""

    def __init__(self, spatial_dims: int, in_channels: int, out_channels: int, kernel_size, stride, padding):
        super().__init__()
        if spatial_dims not in (2, 3):
            raise ValueError("only 2-D and 3-D convolutions are supported")
        self.spatial_dims = spatial_dims
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size = _tup(kernel_size, spatial_dims)
        self.stride = _tup(stride, spatial_dims)
        self.padding = _tup(padding, spatial_dims)
        w = torch.empty(out_channels, in_channels, *self.kernel_size)
        nn.init.kaiming_uniform_(w, a=math.sqrt(5))  # torch's Conv default
        bound = 1 / math.sqrt(in_channels * math.prod(self.kernel_size))
        self.weight = nn.Parameter(w.contiguous(memory_format=_cl_format(spatial_dims)))
        self.bias = nn.Parameter(torch.empty(out_channels).uniform_(-bound, bound))

    def forward(self, x, chan_bias=None, residual=None, gn_groups=0):
        return ops.conv_nd(x, self.weight, self.bias, self.stride, self.padding, chan_bias=chan_bias, residual=residual,
                           gn_groups=gn_groups)

    def extra_repr(self):
        return (f"{self.in_channels}, {self.out_channels}, kernel_size={self.kernel_size}, stride={self.stride}, "
                f"padding={self.padding}")


class ConvTransposeNd(nn.Module):
    """nn.ConvTranspose{2,3}d replacement (weight (Cin, Cout, *k), torch's default initialisation); the kernels are the
    convolution data-gradient kernels (ops.conv_transpose_nd)."""

    def __init__(self, spatial_dims: int, in_channels: int, out_channels: int, kernel_size, stride, padding,
                 output_padding):
        super().__init__()
        if spatial_dims not in (2, 3):
            raise ValueError("only 2-D and 3-D convolutions are supported")
        self.spatial_dims = spatial_dims
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size = _tup(kernel_size, spatial_dims)
        self.stride = _tup(stride, spatial_dims)
        self.padding = _tup(padding, spatial_dims)
        self.output_padding = _tup(output_padding, spatial_dims)
        w = torch.empty(in_channels, out_channels, *self.kernel_size)
        nn.init.kaiming_uniform_(w, a=math.sqrt(5))
        bound = 1 / math.sqrt(out_channels * math.prod(self.kernel_size))   # torch: fan_in of the transposed filter
        self.weight = nn.Parameter(w.contiguous(memory_format=_cl_format(spatial_dims)))
        self.bias = nn.Parameter(torch.empty(out_channels).uniform_(-bound, bound))

    def forward(self, x, chan_bias=None, residual=None, gn_groups=0):
        if chan_bias is not None or residual is not None:
            raise RuntimeError("transposed convolutions have no fused epilogue inputs")
        return ops.conv_transpose_nd(x, self.weight, self.bias, self.stride, self.padding, self.output_padding)

    def extra_repr(self):
        return (f"{self.in_channels}, {self.out_channels}, kernel_size={self.kernel_size}, stride={self.stride}, "
                f"padding={self.padding}, output_padding={self.output_padding}")


class ConvBlock(nn.Module):
    """Stand-in for monai.networks.blocks.Convolution(conv_only=True): single child named `conv`.
    padding=None means same-padding (k-1)//2; is_transposed=True builds a ConvTranspose with MONAI's default
    output_padding = stride - 1."""

    def __init__(self, spatial_dims, in_channels, out_channels, strides=1, kernel_size=3, padding=None,
                 is_transposed=False):
        super().__init__()
        k = _tup(kernel_size, spatial_dims)
        p = tuple((ki - 1) // 2 for ki in k) if padding is None else _tup(padding, spatial_dims)
        s = _tup(strides, spatial_dims)
        if is_transposed:
            self.conv = ConvTransposeNd(spatial_dims, in_channels, out_channels, k, s, p, tuple(si - 1 for si in s))
        else:
            self.conv = ConvNd(spatial_dims, in_channels, out_channels, k, s, p)

    def forward(self, x, chan_bias=None, residual=None, gn_groups=0):
        return self.conv(x, chan_bias=chan_bias, residual=residual, gn_groups=gn_groups)


class GroupNorm(nn.Module):
    def __init__(self, num_groups: int, num_channels: int, eps: float = 1e-5):
        super().__init__()
        if num_channels % num_groups != 0:
            raise ValueError("num_channels must be divisible by num_groups")
        self.num_groups, self.num_channels, self.eps = num_groups, num_channels, eps
        self.weight = nn.Parameter(torch.ones(num_channels))
        self.bias = nn.Parameter(torch.zeros(num_channels))

    def forward(self, x, silu: bool = False, with_skip: bool = False):
        return ops.group_norm(x, self.weight, self.bias, self.num_groups, self.eps, silu=silu, with_skip=with_skip)

    def extra_repr(self):
        return f"{self.num_groups}, {self.num_channels}, eps={self.eps}"


class LayerNorm(nn.Module):
    def __init__(self, dim: int, eps: float = 1e-5):
        super().__init__()
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(dim))
        self.bias = nn.Parameter(torch.zeros(dim))

    def forward(self, x):
        return ops.layer_norm(x, self.weight, self.bias, self.eps)


class Linear(nn.Module):
    def __init__(self, in_features: int, out_features: int, bias: bool = True):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        w = torch.empty(out_features, in_features)
        nn.init.kaiming_uniform_(w, a=math.sqrt(5))
        self.weight = nn.Parameter(w)
        if bias:
            bound = 1 / math.sqrt(in_features)
            self.bias = nn.Parameter(torch.empty(out_features).uniform_(-bound, bound))
        else:
            self.register_parameter("bias", None)

    def forward(self, x):
        return ops.linear(x, self.weight, self.bias)

    def extra_repr(self):
        return f"{self.in_features}, {self.out_features}, bias={self.bias is not None}"


class SiLU(nn.Module):
    def forward(self, x):
        return ops.silu(x)


def zero_module(module: nn.Module) -> nn.Module:
    """unet:62-69."""
    for p in module.parameters():
        p.detach().zero_()
    return module


class SelfAttentionBlock(nn.Module):
    """AttentionBlock of both reference files (unet:345-458, ae:207-323): GroupNorm -> q,k,v Linear
    (+bias) -> per-head softmax(QK^T/sqrt(d))V -> + residual. `proj_attn` is created and stored in the
    state_dict but NEVER applied -- reproduced on purpose."""

    def __init__(self, spatial_dims: int, num_channels: int, num_head_channels=None, norm_num_groups: int = 32,
                 norm_eps: float = 1e-6, use_flash_attention: bool = False):
        super().__init__()
        self.use_flash_attention = use_flash_attention  # accepted; the native kernels are always used
        self.spatial_dims = spatial_dims
        self.num_channels = num_channels
        self.num_heads = num_channels // num_head_channels if num_head_channels is not None else 1
        self.scale = 1 / math.sqrt(num_channels / self.num_heads)
        self.norm = GroupNorm(norm_num_groups, num_channels, norm_eps)
        self.to_q = Linear(num_channels, num_channels)
        self.to_k = Linear(num_channels, num_channels)
        self.to_v = Linear(num_channels, num_channels)
        self.proj_attn = Linear(num_channels, num_channels)

    def _fused_qkv(self):
        """ops.FusedLinearParams of to_q/to_k/to_v when a FlatAdamW owns them (cached per storage), else None."""
        wq = self.to_q.weight
        if getattr(wq, "_mig_flat", None) is None:
            return None
        key = (wq.data_ptr(), self.to_k.weight.data_ptr(), self.to_v.weight.data_ptr(), id(wq._mig_flat))
        cache = self.__dict__.get("_mig_qkv")
        if cache is None or cache[0] != key:
            cache = (key, ops.fuse_linears([("q", self.to_q), ("k", self.to_k), ("v", self.to_v)]))
            self.__dict__["_mig_qkv"] = cache
        return cache[1]

    def _eval_qkv(self):
        """(3C, C) bf16 weight and (3C,) fp32 bias of to_q / to_k / to_v stacked, cached until one of them changes."""
        ps = (self.to_q.weight, self.to_k.weight, self.to_v.weight, self.to_q.bias, self.to_k.bias, self.to_v.bias)
        key = tuple((p.data_ptr(), p._version) for p in ps)
        cache = self.__dict__.get("_mig_qkv_eval")
        if cache is None or cache[0] != key:
            with torch.no_grad():
                w = torch.cat([p.detach().to(torch.bfloat16) for p in ps[:3]], 0).contiguous()
                b = torch.cat([p.detach().float() for p in ps[3:]], 0).contiguous()
            cache = (key, w, b)
            self.__dict__["_mig_qkv_eval"] = cache
        return cache[1], cache[2]

    def forward(self, x):
        x = ops.to_channels_last(x, x.dtype if x.dtype in (torch.float32, torch.bfloat16) else torch.float32)
        B, Cc = x.shape[0], x.shape[1]
        h, x = self.norm(x, with_skip=True)   # x also feeds the residual add below: its gradient joins inside norm's backward
        # channels-last memory IS the (B, L, C) token matrix in the reference's d,h,w order (unet:428-434)
        tokens = h.permute(0, *range(2, h.ndim), 1).reshape(B, -1, Cc)
        fused = self._fused_qkv()
        if fused is None and not torch.is_grad_enabled() and tokens.dtype == torch.bfloat16:
            # inference without an optimiser (sampling): the same single projection on a cached concatenation of the three
            # weight matrices; the attention kernel reads q / k / v as column blocks of its output (6 x (3 GEMMs -> 1)
            # per reverse step of the config-4 U-Net)
            w, b = self._eval_qkv()
            qkv = ops.linear(tokens, w, b)
            o = ops.sdpa_qkv(qkv, self.num_heads, self.scale)
        elif fused is not None:
            # to_q / to_k / to_v (unet:436-438) as ONE projection GEMM: the flat optimiser keeps the three weight matrices
            # side by side; attention reads the column slices in place and returns ONE gradient tensor
            fused.refresh()
            qkv = ops.linear(tokens, fused.weight, fused.bias)
            o = ops.sdpa_qkv(qkv, self.num_heads, self.scale, order=fused.order)
        else:
            q, k, v = self.to_q(tokens), self.to_k(tokens), self.to_v(tokens)
            o = ops.sdpa(q, k, v, self.num_heads, self.scale)
        o = o.reshape(B, *x.shape[2:], Cc).permute(0, x.ndim - 1, *range(1, x.ndim - 1))
        return ops.add(o, x)
