"""Deterministic parameter/input synthesis shared by oracle/gen_golden.py and tests/.

TEST INFRASTRUCTURE. Golden files stay small because parameters are not stored: they are
re-drawn from a seeded CPU generator, in sorted key order, from the (key -> shape) table the golden
holds. zero_module()'d parameters (unet:62-69) are thereby re-randomised, as parity on an
all-zero output would be vacuous.
"""
from __future__ import annotations

import math

import torch

CASES = {
    # LDM-default structure at reduced width (create_ddpm_dict, configuration.py:865-902)
    "unet3d_small": dict(kind="unet", batch=2, in_shape=(3, 8, 8, 8), cfg=dict(
        spatial_dims=3, in_channels=3, out_channels=3, num_res_blocks=2, num_channels=[32, 64, 96],
        attention_levels=[False, True, True], num_head_channels=[0, 64, 96], norm_num_groups=16,
        strides=[[1, 1, 1], [2, 2, 2], [2, 2, 2]], kernel_sizes=[[3, 3, 3]] * 3, paddings=[[1, 1, 1]] * 3)),
    # BASELINE config 3 at FULL width (create_ddpm_dict: 256/512/768, one 512-/768-channel head per attention level,
    # 441 M parameters; K up to 41 472, 1536-channel skip-concat inputs) on a small latent so the CPU reference finishes
    # in seconds. Parameters are re-drawn from the seed (never stored), so the golden stays ~1.5 MB.
    "unet3d_ldm_width": dict(kind="unet", batch=2, in_shape=(3, 8, 8, 8), cfg=dict(
        spatial_dims=3, in_channels=3, out_channels=3, num_res_blocks=2, num_channels=[256, 512, 768],
        attention_levels=[False, True, True], num_head_channels=[0, 512, 768], norm_num_groups=32,
        strides=[[1, 1, 1], [2, 2, 2], [2, 2, 2]], kernel_sizes=[[3, 3, 3]] * 3, paddings=[[1, 1, 1]] * 3)),
    # anisotropic strides, multi-head attention, odd spatial sizes (BASELINE config 4 shape class)
    "unet3d_aniso": dict(kind="unet", batch=1, in_shape=(1, 12, 12, 6), cfg=dict(
        spatial_dims=3, in_channels=1, out_channels=1, num_res_blocks=1, num_channels=[16, 32, 64],
        attention_levels=[False, False, True], num_head_channels=[0, 0, 16], norm_num_groups=8,
        strides=[[1, 1, 1], [2, 2, 1], [2, 2, 2]], kernel_sizes=[[3, 3, 3]] * 3, paddings=[[1, 1, 1]] * 3)),
    # BASELINE config 1 shape class: 2D DDPM
    "unet2d_small": dict(kind="unet", batch=2, in_shape=(1, 16, 16), cfg=dict(
        spatial_dims=2, in_channels=1, out_channels=1, num_res_blocks=2, num_channels=[32, 64, 64],
        attention_levels=[False, True, True], num_head_channels=[0, 64, 64], norm_num_groups=32,
        strides=[[1, 1], [2, 2], [2, 2]], kernel_sizes=[[3, 3]] * 3, paddings=[[1, 1]] * 3)),
    # conditioning path (SpatialTransformer + class embedding), "next" row f2
    "unet3d_cond": dict(kind="unet", batch=2, in_shape=(2, 4, 4, 4), context=(3, 24), classes=5, cfg=dict(
        spatial_dims=3, in_channels=2, out_channels=2, num_res_blocks=1, num_channels=[32, 64],
        attention_levels=[False, True], num_head_channels=[0, 16], norm_num_groups=16,
        with_conditioning=True, cross_attention_dim=24, num_class_embeds=5, transformer_num_layers=1,
        strides=[[1, 1, 1], [2, 2, 2]], kernel_sizes=[[3, 3, 3]] * 2, paddings=[[1, 1, 1]] * 2)),
    # resblock_updown=True (unet:641-644,757-768,1231-1242): AvgPool / nearest ResnetBlocks instead of strided convs.
    # Only consistent when the level's kernel equals its stride (the pool has no padding); hand-written, never planned.
    "unet3d_updown": dict(kind="unet", batch=2, in_shape=(2, 8, 8, 4), cfg=dict(
        spatial_dims=3, in_channels=2, out_channels=2, num_res_blocks=1, num_channels=[16, 32],
        attention_levels=[False, True], num_head_channels=[0, 16], norm_num_groups=8, resblock_updown=True,
        strides=[[1, 1, 1], [2, 2, 1]], kernel_sizes=[[3, 3, 3], [2, 2, 1]], paddings=[[1, 1, 1], [0, 0, 0]])),
    # AE-default structure at reduced width (create_autoencoder_dict, configuration.py:821-862)
    "ae3d_small": dict(kind="ae", batch=2, in_shape=(1, 16, 16, 16), cfg=dict(
        spatial_dims=3, in_channels=1, out_channels=1, num_res_blocks=2, num_channels=[16, 32, 64],
        attention_levels=[False, False, False], latent_channels=3, norm_num_groups=16,
        with_encoder_nonlocal_attn=False, with_decoder_nonlocal_attn=False,
        downsample_parameters=[[[1, 1, 1], [3, 3, 3], [1, 1, 1]], [[2, 2, 2], [3, 3, 3], [1, 1, 1]],
                               [[2, 2, 2], [3, 3, 3], [1, 1, 1]]],
        upsample_parameters=[[[2, 2, 2], [3, 3, 3], [1, 1, 1]], [[2, 2, 2], [3, 3, 3], [1, 1, 1]]])),
    # AE with attention + non-local blocks, anisotropic thin axis (kernel 1 / pad 0 / stride 1 on it)
    "ae3d_attn_aniso": dict(kind="ae", batch=1, in_shape=(2, 16, 16, 4), cfg=dict(
        spatial_dims=3, in_channels=2, out_channels=2, num_res_blocks=1, num_channels=[16, 32],
        attention_levels=[False, True], latent_channels=4, norm_num_groups=8,
        with_encoder_nonlocal_attn=True, with_decoder_nonlocal_attn=True,
        downsample_parameters=[[[1, 1, 1], [3, 3, 1], [1, 1, 0]], [[2, 2, 1], [3, 3, 1], [1, 1, 0]]],
        upsample_parameters=[[[2, 2, 1], [3, 3, 1], [1, 1, 0]]])),
    # transposed-convolution upsampling (ae:66-76; never emitted by the planner, configuration.py:843)
    "ae3d_convtranspose": dict(kind="ae", batch=1, in_shape=(1, 16, 16, 8), cfg=dict(
        spatial_dims=3, in_channels=1, out_channels=1, num_res_blocks=1, num_channels=[16, 32],
        attention_levels=[False, False], latent_channels=3, norm_num_groups=8,
        with_encoder_nonlocal_attn=False, with_decoder_nonlocal_attn=False, use_convtranspose=True,
        downsample_parameters=[[[1, 1, 1], [3, 3, 3], [1, 1, 1]], [[2, 2, 1], [3, 3, 3], [1, 1, 1]]],
        upsample_parameters=[[[2, 2, 1], [3, 3, 3], [1, 1, 1]]])),
    "ae2d_small": dict(kind="ae", batch=2, in_shape=(1, 32, 32), cfg=dict(
        spatial_dims=2, in_channels=1, out_channels=1, num_res_blocks=1, num_channels=[32, 64],
        attention_levels=[False, False], latent_channels=3, norm_num_groups=16,
        with_encoder_nonlocal_attn=False, with_decoder_nonlocal_attn=False,
        downsample_parameters=[[[1, 1], [3, 3], [1, 1]], [[2, 2], [3, 3], [1, 1]]],
        upsample_parameters=[[[2, 2], [3, 3], [1, 1]]])),
}


def golden_params(shapes: dict, seed: int) -> dict:
    """name -> fp32 tensor; scale ~ 1/sqrt(fan_in) for matrices/filters, ~N(1,.1) norm gains, N(0,.1) biases."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for name in sorted(shapes):
        shape = tuple(shapes[name])
        r = torch.randn(shape, generator=g)
        if len(shape) >= 2:
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            out[name] = r / math.sqrt(fan_in)
        elif name.endswith("weight"):  # 1-D weight == a norm gain
            out[name] = 1.0 + 0.1 * r
        else:
            out[name] = 0.1 * r
    return out


def golden_inputs(case: dict, seed: int) -> dict:
    g = torch.Generator().manual_seed(seed + 7919)
    B = case["batch"]
    d = {"x": torch.randn((B, *case["in_shape"]), generator=g)}
    if case["kind"] == "unet":
        d["timesteps"] = torch.randint(0, 1000, (B,), generator=g)
        d["probe"] = torch.randn((B, case["cfg"]["out_channels"], *case["in_shape"][1:]), generator=g)
        if "context" in case:
            d["context"] = torch.randn((B, *case["context"]), generator=g)
            d["class_labels"] = torch.randint(0, case["classes"], (B,), generator=g)
    else:
        d["x"] = torch.rand((B, *case["in_shape"]), generator=g)  # loader clamps images to [0,1]
        d["eps"] = None  # filled by the generator once the latent shape is known
    return d


def sketch(t: torch.Tensor, n: int = 64) -> torch.Tensor:
    """Small fingerprint of a tensor: [sum, abs-sum, sq-sum, n strided samples]."""
    f = t.detach().double().flatten()
    step = max(1, f.numel() // n)
    return torch.cat([torch.stack([f.sum(), f.abs().sum(), (f * f).sum()]), f[::step][:n]]).float()
