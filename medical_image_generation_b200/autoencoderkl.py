"""B200-native drop-in for medimgen's strided `AutoencoderKL` (medimgen/autoencoderkl_with_strides.py, `ae:N`).

Same constructor, `encode / sampling / decode / reconstruct / forward / encode_stage_2_inputs /
decode_stage_2_outputs` methods and `state_dict` layout (`encoder.blocks.{i}...`, `decoder.blocks.{i}...`,
`quant_conv_mu`, `quant_conv_log_sigma`, `post_quant_conv`). All compute runs through ops.py (sm_100a kernels).
"""
from __future__ import annotations

from typing import Sequence

import torch
from torch import nn

from . import ops
from .layers import ConvBlock, GroupNorm, SelfAttentionBlock, _tup

__all__ = ["AutoencoderKL"]

_COMPUTE = (torch.float32, torch.bfloat16)


def _entry(x):
    return ops.to_channels_last(x, x.dtype if x.dtype in _COMPUTE else torch.float32)


class Upsample(nn.Module):
    """ae:52-106: nearest x stride then 3^n conv (pad 1), or a transposed conv when use_convtranspose."""

    def __init__(self, spatial_dims, in_channels, use_convtranspose, stride=2, kernel_size=4, padding=1):
        super().__init__()
        self.stride = stride
        if use_convtranspose:   # ae:66-76
            self.conv = ConvBlock(spatial_dims, in_channels, in_channels, strides=stride, kernel_size=kernel_size,
                                  padding=padding, is_transposed=True)
        else:
            self.conv = ConvBlock(spatial_dims, in_channels, in_channels, strides=1, kernel_size=3, padding=1)
        self.use_convtranspose = use_convtranspose

    def forward(self, x):
        if self.use_convtranspose:
            return self.conv(_entry(x))
        c = self.conv.conv
        return ops.upsample_conv_nd(_entry(x), c.weight, c.bias, _tup(self.stride, x.ndim - 2), c.padding)


class Downsample(nn.Module):
    """ae:109-133: strided conv."""

    def __init__(self, spatial_dims, in_channels, stride=2, kernel_size=4, padding=1):
        super().__init__()
        self.conv = ConvBlock(spatial_dims, in_channels, in_channels, strides=stride, kernel_size=kernel_size,
                              padding=padding)

    def forward(self, x):
        return self.conv(_entry(x))


class ResBlock(nn.Module):
    """ae:136-204: GN+SiLU -> conv -> GN+SiLU -> conv (+ shortcut fused into the second conv's epilogue)."""

    def __init__(self, spatial_dims, in_channels, norm_num_groups, norm_eps, out_channels):
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = in_channels if out_channels is None else out_channels
        self.norm1 = GroupNorm(norm_num_groups, in_channels, norm_eps)
        self.conv1 = ConvBlock(spatial_dims, self.in_channels, self.out_channels, strides=1, kernel_size=3, padding=1)
        self.norm2 = GroupNorm(norm_num_groups, out_channels, norm_eps)
        self.conv2 = ConvBlock(spatial_dims, self.out_channels, self.out_channels, strides=1, kernel_size=3, padding=1)
        if self.in_channels != self.out_channels:
            self.nin_shortcut = ConvBlock(spatial_dims, self.in_channels, self.out_channels, strides=1, kernel_size=1,
                                          padding=0)
        else:
            self.nin_shortcut = nn.Identity()

    def forward(self, x):
        x = _entry(x)
        h, x = self.norm1(x, silu=True, with_skip=True)   # the skip gradient is added inside norm1's backward kernel
        h = self.conv1(h, gn_groups=self.norm2.num_groups)
        h._mig_sole_consumer_gn = True   # norm2 is conv1's only consumer: its backward hands conv1 the column sums of dy
        h = self.norm2(h, silu=True)
        skip = x if self.in_channels == self.out_channels else self.nin_shortcut(x)
        return self.conv2(h, residual=skip, gn_groups=self.norm2.num_groups)


class AttentionBlock(SelfAttentionBlock):
    """ae:207-323 (num_head_channels=None -> one head of C channels)."""


class Encoder(nn.Module):
    """ae:326-470: flat `blocks` list; the bare GroupNorm before the last conv has NO activation."""

    def __init__(self, spatial_dims, in_channels, num_channels, out_channels, num_res_blocks, norm_num_groups,
                 norm_eps, attention_levels, with_nonlocal_attn=True, use_flash_attention=False, strides=2,
                 kernel_sizes=4, paddings=1):
        super().__init__()
        self.spatial_dims, self.in_channels, self.num_channels = spatial_dims, in_channels, num_channels
        self.out_channels, self.num_res_blocks = out_channels, num_res_blocks
        self.norm_num_groups, self.norm_eps, self.attention_levels = norm_num_groups, norm_eps, attention_levels

        def lvl(v, i):
            return v if isinstance(v, int) else v[i]

        blocks: list = [ConvBlock(spatial_dims, in_channels, num_channels[0], strides=lvl(strides, 0),
                                  kernel_size=lvl(kernel_sizes, 0), padding=lvl(paddings, 0))]
        output_channel = num_channels[0]
        L = len(num_channels)
        for i in range(L):
            input_channel, output_channel = output_channel, num_channels[i]
            for _ in range(self.num_res_blocks[i]):
                blocks.append(ResBlock(spatial_dims, input_channel, norm_num_groups, norm_eps, output_channel))
                input_channel = output_channel
                if attention_levels[i]:
                    blocks.append(AttentionBlock(spatial_dims, input_channel, None, norm_num_groups, norm_eps,
                                                 use_flash_attention))
            if i != L - 1:
                blocks.append(Downsample(spatial_dims, input_channel, stride=lvl(strides, i + 1),
                                         kernel_size=lvl(kernel_sizes, i + 1), padding=lvl(paddings, i + 1)))
        if with_nonlocal_attn is True:
            c = num_channels[-1]
            blocks.append(ResBlock(spatial_dims, c, norm_num_groups, norm_eps, c))
            blocks.append(AttentionBlock(spatial_dims, c, None, norm_num_groups, norm_eps, use_flash_attention))
            blocks.append(ResBlock(spatial_dims, c, norm_num_groups, norm_eps, c))
        blocks.append(GroupNorm(norm_num_groups, num_channels[-1], norm_eps))
        blocks.append(ConvBlock(spatial_dims, num_channels[-1], out_channels, strides=1, kernel_size=3, padding=1))
        self.blocks = nn.ModuleList(blocks)

    def forward(self, x):
        x = _entry(x)
        for block in self.blocks:
            x = block(x)
        return x


class Decoder(nn.Module):
    """ae:473-622."""

    def __init__(self, spatial_dims, num_channels, in_channels, out_channels, num_res_blocks, norm_num_groups,
                 norm_eps, attention_levels, with_nonlocal_attn=True, use_flash_attention=False,
                 use_convtranspose=False, strides=2, kernel_sizes=4, paddings=1):
        super().__init__()
        self.spatial_dims, self.num_channels, self.in_channels = spatial_dims, num_channels, in_channels
        self.out_channels, self.num_res_blocks = out_channels, num_res_blocks
        self.norm_num_groups, self.norm_eps, self.attention_levels = norm_num_groups, norm_eps, attention_levels

        def lvl(v, i):
            return v if isinstance(v, int) else v[i]

        rch = list(reversed(num_channels))
        blocks: list = [ConvBlock(spatial_dims, in_channels, rch[0], strides=1, kernel_size=3, padding=1)]
        if with_nonlocal_attn is True:
            blocks.append(ResBlock(spatial_dims, rch[0], norm_num_groups, norm_eps, rch[0]))
            blocks.append(AttentionBlock(spatial_dims, rch[0], None, norm_num_groups, norm_eps, use_flash_attention))
            blocks.append(ResBlock(spatial_dims, rch[0], norm_num_groups, norm_eps, rch[0]))
        rattn, rres = list(reversed(attention_levels)), list(reversed(num_res_blocks))
        block_out = rch[0]
        L = len(rch)
        for i in range(L):
            block_in, block_out = block_out, rch[i]
            for _ in range(rres[i]):
                blocks.append(ResBlock(spatial_dims, block_in, norm_num_groups, norm_eps, block_out))
                block_in = block_out
                if rattn[i]:
                    blocks.append(AttentionBlock(spatial_dims, block_in, None, norm_num_groups, norm_eps,
                                                 use_flash_attention))
            if i != L - 1:
                blocks.append(Upsample(spatial_dims, block_in, use_convtranspose, stride=lvl(strides, i),
                                       kernel_size=lvl(kernel_sizes, i), padding=lvl(paddings, i)))
        blocks.append(GroupNorm(norm_num_groups, block_in, norm_eps))
        blocks.append(ConvBlock(spatial_dims, block_in, out_channels, strides=1, kernel_size=3, padding=1))
        self.blocks = nn.ModuleList(blocks)

    def forward(self, x):
        x = _entry(x)
        for block in self.blocks:
            x = block(x)
        return x


class AutoencoderKL(nn.Module):
    """Drop-in for ae:625-838. Extra optional keyword `compute_dtype` (bf16 default, fp32 for the 1e-4 bar)."""

    def __init__(self, spatial_dims: int, in_channels: int = 1, out_channels: int = 1,
                 num_res_blocks: Sequence[int] | int = (2, 2, 2, 2), num_channels: Sequence[int] = (32, 64, 64, 64),
                 attention_levels: Sequence[bool] = (False, False, True, True), latent_channels: int = 3,
                 norm_num_groups: int = 32, norm_eps: float = 1e-6, with_encoder_nonlocal_attn: bool = True,
                 with_decoder_nonlocal_attn: bool = True, use_flash_attention: bool = False,
                 use_checkpointing: bool = False, use_convtranspose: bool = False,
                 downsample_parameters=((2, 4, 1), (2, 4, 1), (2, 4, 1)),
                 upsample_parameters=((2, 4, 1), (2, 4, 1), (2, 4, 1)),
                 compute_dtype: torch.dtype = torch.bfloat16) -> None:
        super().__init__()
        if any((c % norm_num_groups) != 0 for c in num_channels):
            raise ValueError("AutoencoderKL expects all num_channels being multiple of norm_num_groups")
        if len(num_channels) != len(attention_levels):
            raise ValueError("AutoencoderKL expects num_channels being same size of attention_levels")
        if isinstance(num_res_blocks, int):
            num_res_blocks = (num_res_blocks,) * len(num_channels)
        if len(num_res_blocks) != len(num_channels):
            raise ValueError("`num_res_blocks` should be a single integer or a tuple of integers with the same length "
                             "as `num_channels`.")
        self.encoder = Encoder(spatial_dims, in_channels, num_channels, latent_channels, num_res_blocks,
                               norm_num_groups, norm_eps, attention_levels, with_encoder_nonlocal_attn,
                               use_flash_attention, strides=[p[0] for p in downsample_parameters],
                               kernel_sizes=[p[1] for p in downsample_parameters],
                               paddings=[p[2] for p in downsample_parameters])
        self.decoder = Decoder(spatial_dims, num_channels, latent_channels, out_channels, num_res_blocks,
                               norm_num_groups, norm_eps, attention_levels, with_decoder_nonlocal_attn,
                               use_flash_attention, use_convtranspose, strides=[p[0] for p in upsample_parameters],
                               kernel_sizes=[p[1] for p in upsample_parameters],
                               paddings=[p[2] for p in upsample_parameters])
        self.quant_conv_mu = ConvBlock(spatial_dims, latent_channels, latent_channels, strides=1, kernel_size=1,
                                       padding=0)
        self.quant_conv_log_sigma = ConvBlock(spatial_dims, latent_channels, latent_channels, strides=1,
                                              kernel_size=1, padding=0)
        self.post_quant_conv = ConvBlock(spatial_dims, latent_channels, latent_channels, strides=1, kernel_size=1,
                                         padding=0)
        self.latent_channels = latent_channels
        self.use_checkpointing = use_checkpointing
        self.compute_dtype = compute_dtype

    def _out_dtype(self, x):
        return x.dtype if x.dtype.is_floating_point else torch.float32

    def encode(self, x):
        """ae:753-771 -> (z_mu, z_sigma) as standard (N,C,*sp) tensors of x's dtype."""
        od = self._out_dtype(x)
        xin = ops.to_channels_last(x, self.compute_dtype)
        if self.use_checkpointing:
            h = torch.utils.checkpoint.checkpoint(self.encoder, xin, use_reentrant=False)
        else:
            h = self.encoder(xin)
        z_mu = self.quant_conv_mu(h)
        z_sigma = ops.vae_sigma(self.quant_conv_log_sigma(h))
        return ops.from_channels_last(z_mu, od), ops.from_channels_last(z_sigma, od)

    def sampling(self, z_mu, z_sigma):
        """ae:773-788: z = mu + eps*sigma with eps ~ N(0,1) drawn like torch.randn_like(z_sigma)."""
        eps = torch.randn_like(z_sigma)
        return ops.vae_reparam(z_mu, z_sigma, eps)

    def reconstruct(self, x):
        z_mu, _ = self.encode(x)
        return self.decode(z_mu)

    def decode(self, z):
        """ae:804-819."""
        od = self._out_dtype(z)
        h = self.post_quant_conv(ops.to_channels_last(z, self.compute_dtype))
        if self.use_checkpointing:
            dec = torch.utils.checkpoint.checkpoint(self.decoder, h, use_reentrant=False)
        else:
            dec = self.decoder(h)
        return ops.from_channels_last(dec, od)

    def forward(self, x):
        z_mu, z_sigma = self.encode(x)
        z = self.sampling(z_mu, z_sigma)
        return self.decode(z), z_mu, z_sigma

    def encode_stage_2_inputs(self, x):
        z_mu, z_sigma = self.encode(x)
        return self.sampling(z_mu, z_sigma)

    def decode_stage_2_outputs(self, z):
        return self.decode(z)

    @staticmethod
    def initialize(module):
        """ae:836-838 (InitWeights_He, ae:41-49): kaiming-normal filters, zero biases; applied with .apply()."""
        from .layers import ConvNd, ConvTransposeNd
        if isinstance(module, (ConvNd, ConvTransposeNd)):   # ae:46 also covers ConvTranspose{2,3}d
            with torch.no_grad():
                w = torch.empty(module.weight.shape)
                nn.init.kaiming_normal_(w, a=1e-2)
                module.weight.copy_(w)
                module.bias.zero_()
