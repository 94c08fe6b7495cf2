"""generative.networks.nets as far as the reference trainers import it.

* VQVAE (train_ldm.py:30, train_autoencoder.py:26; optional `-l vq` latent space): placeholder, not part of BASELINE.json's
  north star (the trainers only use the name in `isinstance` checks unless `-l vq` is chosen).
* PatchDiscriminator (train_autoencoder.py:26,600): the PatchGAN discriminator of the adversarial AE loss.
  monai-generative is not installed in this image and not under /root/reference, so this is a restatement of its published
  architecture [upstream-memory; parity unpinned]: Conv(k=4, s=2, bias) + LeakyReLU(0.2), then `num_layers_d` blocks of
  Conv(k=4, s=2 -- the last one s=1 --, no bias) + BatchNorm + LeakyReLU(0.2) with doubling widths, then a one-channel
  Conv(k=4, s=1, bias); weights N(0, 0.02), BatchNorm gains N(1, 0.02); `forward` returns the list of all intermediate
  feature maps (the trainer takes `[-1]`, train_autoencoder.py:380-382,420). Module names follow MONAI's `Convolution`
  block (`<name>.conv`, `<name>.adn.N`, `<name>.adn.A`) so `discriminator_state_dict` checkpoints keep their key layout.
  The discriminator is OUTSIDE the B200 hot path (BASELINE.json north star: U-Net / AutoencoderKL blocks + DDPM loop): it is
  plain torch.nn, a few strided 4^3 convolutions on the image, and is here so that the unmodified
  `train_autoencoder.AutoEncoder` can run end to end against the B200 AutoencoderKL."""
from __future__ import annotations

import torch
from torch import nn

from .._placeholder import placeholder

VQVAE = placeholder("networks.nets.VQVAE", "the VQ-VAE latent space is not part of BASELINE.json's north star")

_CONV = {2: nn.Conv2d, 3: nn.Conv3d}
_BN = {2: nn.BatchNorm2d, 3: nn.BatchNorm3d}


def _block(spatial_dims, cin, cout, kernel_size, stride, padding, bias, norm, act_slope, dropout):
    """MONAI `Convolution(..., conv_only=False)`: Sequential(conv, adn=Sequential(N, [D], A))."""
    blk = nn.Sequential()
    blk.add_module("conv", _CONV[spatial_dims](cin, cout, kernel_size, stride, padding, bias=bias))
    adn = nn.Sequential()
    if norm:
        adn.add_module("N", _BN[spatial_dims](cout))
    if dropout:
        adn.add_module("D", nn.Dropout(dropout))
    adn.add_module("A", nn.LeakyReLU(negative_slope=act_slope))
    blk.add_module("adn", adn)
    return blk


class PatchDiscriminator(nn.Sequential):
    def __init__(self, spatial_dims: int, num_channels: int, in_channels: int, out_channels: int = 1,
                 num_layers_d: int = 3, kernel_size: int = 4, activation=("LEAKYRELU", {"negative_slope": 0.2}),
                 norm="BATCH", bias: bool = False, padding=1, dropout=0.0, last_conv_kernel_size=None) -> None:
        super().__init__()
        if spatial_dims not in (2, 3):
            raise ValueError("PatchDiscriminator: spatial_dims must be 2 or 3")
        if str(norm).upper() != "BATCH":
            raise NotImplementedError("PatchDiscriminator shim: only norm='BATCH' (the reference's setting) is provided")
        slope = 0.2
        if isinstance(activation, (tuple, list)) and len(activation) > 1:
            slope = float(activation[1].get("negative_slope", 0.2))
        self.num_layers_d, self.num_channels = num_layers_d, num_channels
        if last_conv_kernel_size is None:
            last_conv_kernel_size = kernel_size
        self.add_module("initial_conv", _block(spatial_dims, in_channels, num_channels, kernel_size, 2, padding, True, False,
                                               slope, dropout))
        cin, cout = num_channels, num_channels * 2
        for l_ in range(num_layers_d):
            stride = 1 if l_ == num_layers_d - 1 else 2
            self.add_module("%d" % l_, _block(spatial_dims, cin, cout, kernel_size, stride, padding, bias, True, slope, dropout))
            cin, cout = cout, cout * 2
        final = nn.Sequential()
        final.add_module("conv", _CONV[spatial_dims](cin, out_channels, last_conv_kernel_size, 1,
                                                     int((last_conv_kernel_size - 1) / 2), bias=True))
        self.add_module("final_conv", final)
        self.apply(self.initialise_weights)

    def forward(self, x: torch.Tensor) -> list[torch.Tensor]:
        out = [x]
        for submodel in self.children():
            out.append(submodel(out[-1]))
        return out[1:]

    @staticmethod
    def initialise_weights(m: nn.Module) -> None:
        if isinstance(m, (nn.Conv2d, nn.Conv3d)):
            nn.init.normal_(m.weight.data, 0.0, 0.02)
        elif isinstance(m, (nn.BatchNorm2d, nn.BatchNorm3d)):
            nn.init.normal_(m.weight.data, 1.0, 0.02)
            nn.init.constant_(m.bias.data, 0)
