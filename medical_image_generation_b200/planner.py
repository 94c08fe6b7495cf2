"""Shape planner: the pure functions of medimgen/configuration.py:751-902 that turn a patch size into the per-level
stride / kernel / padding lists and the constructor kwargs of the two networks. They define the benchmark shapes
(SURVEY.md section 8d); the dataset-fingerprint / IO half of the reference planner is out of scope.
"""
from __future__ import annotations

from typing import Sequence


def compute_downsample_parameters(input_size: Sequence[int], num_layers: int):
    """configuration.py:751-797. An axis that is at most half of the longest other axis is 'thin': it gets
    kernel 1 / padding 0 and is not strided; every other axis gets kernel 3 / padding 1 and (from the second
    layer on) stride 2. Returns [[stride], [kernel], [padding]] per layer."""
    size = [int(v) for v in input_size]
    nd = len(size)
    layers = []
    for layer in range(num_layers):
        stride, kernel, padding = [], [], []
        for axis in range(nd):
            rest = [size[j] for j in range(nd) if j != axis]
            thin = size[axis] <= 0.5 * (max(rest) if rest else size[axis])
            kernel.append(1 if thin else 3)
            padding.append(0 if thin else 1)
            stride.append(1 if (thin or layer == 0) else 2)
        if layer > 0:
            size = [(size[a] + 2 * padding[a] - kernel[a]) // stride[a] + 1 for a in range(nd)]
        layers.append([stride, kernel, padding])
    return layers


def compute_output_size(input_size: Sequence[int], downsample_parameters):
    """configuration.py:800-818."""
    size = [int(v) for v in input_size]
    for stride, kernel, padding in downsample_parameters:
        size = [(size[a] + 2 * padding[a] - kernel[a]) // stride[a] + 1 for a in range(len(size))]
    return size


def _vae_levels(patch_size) -> int:
    longest = max(patch_size)
    return 1 if longest <= 96 else (2 if longest <= 384 else 3)


def autoencoder_kwargs(patch_size, in_channels: int = 1, latent_channels: int = 8, levels: int | None = None) -> dict:
    """create_autoencoder_dict (configuration.py:821-862) for an already-chosen patch size."""
    nd = len(patch_size)
    n = _vae_levels(patch_size) if levels is None else levels
    base = [64, 128, 256, 256] if nd == 2 else [32, 64, 128, 128]
    down = compute_downsample_parameters(patch_size, n + 1)
    return dict(spatial_dims=nd, in_channels=in_channels, out_channels=in_channels, latent_channels=latent_channels,
                num_res_blocks=2, with_encoder_nonlocal_attn=False, with_decoder_nonlocal_attn=False,
                use_flash_attention=False, use_checkpointing=False, use_convtranspose=False,
                num_channels=base[:n + 1], attention_levels=[False] * (n + 1), norm_num_groups=16,
                downsample_parameters=down, upsample_parameters=list(reversed(down))[:-1])


def ddpm_kwargs(latent_size, latent_channels: int = 8) -> dict:
    """create_ddpm_dict (configuration.py:865-902) for a latent of `latent_size`."""
    p = compute_downsample_parameters(latent_size, 3)
    return dict(spatial_dims=len(latent_size), in_channels=latent_channels, out_channels=latent_channels,
                num_res_blocks=2, use_flash_attention=False, num_channels=[256, 512, 768],
                attention_levels=[False, True, True], num_head_channels=[0, 512, 768],
                strides=[q[0] for q in p], kernel_sizes=[q[1] for q in p], paddings=[q[2] for q in p])


LDM_SCHEDULER_KWARGS = dict(num_train_timesteps=1000, schedule="scaled_linear_beta", beta_start=0.0015,
                            beta_end=0.0205, prediction_type="epsilon")  # configuration.py:1012-1013
