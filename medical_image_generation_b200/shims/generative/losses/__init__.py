"""generative.losses (train_autoencoder.py:27).

* PatchAdversarialLoss (train_autoencoder.py:47,381-385,421): restated from monai-generative's published definition
  [upstream-memory; parity unpinned -- the package is neither installed nor under /root/reference]. `least_squares` (the
  reference's criterion) = MSE between LeakyReLU(0.05)(logits) and a constant 1 / 0 target; `bce` and `hinge` follow the
  same upstream conventions. Accepts one logits tensor or a list (multi-scale discriminators) and averages over it.
* PerceptualLoss: LPIPS-VGG with (fake-3D) slice sampling, weights from a local file -- see .perceptual."""
from __future__ import annotations

import warnings

import torch
from torch import nn

from .perceptual import PerceptualLoss  # noqa: F401


class PatchAdversarialLoss(nn.Module):
    def __init__(self, reduction: str = "mean", criterion: str = "least_squares", no_activation_leastsq: bool = False) -> None:
        super().__init__()
        criterion = str(criterion).lower()
        if criterion not in ("bce", "hinge", "least_squares"):
            raise ValueError("Unrecognised criterion entered for Adversarial Loss. Must be one in: bce, hinge, least_squares")
        self.real_label, self.fake_label = 1.0, 0.0
        self.criterion, self.reduction = criterion, reduction
        self.activation = None
        if criterion == "bce":
            self.activation = nn.Sigmoid()
            self.loss_fct = nn.BCELoss(reduction=reduction)
        elif criterion == "hinge":
            self.activation = nn.Tanh()
            self.fake_label = -1.0
            self.loss_fct = None
        else:
            if not no_activation_leastsq:
                self.activation = nn.LeakyReLU(negative_slope=0.05)
            self.loss_fct = nn.MSELoss(reduction=reduction)

    def get_target_tensor(self, x: torch.Tensor, target_is_real: bool) -> torch.Tensor:
        return torch.full_like(x, self.real_label if target_is_real else self.fake_label, requires_grad=False)

    def forward(self, input, target_is_real: bool, for_discriminator: bool):
        if not for_discriminator and not target_is_real:
            target_is_real = True   # the generator always wants its fakes judged real
            warnings.warn("Variable target_is_real has been set to False, but for_discriminator is set to False. "
                          "To optimise a generator, target_is_real must be set to True.")
        outs = input if isinstance(input, (list, tuple)) else [input]
        losses = []
        for d in outs:
            if self.activation is not None:
                d = self.activation(d)
            if self.criterion == "hinge":
                if for_discriminator:
                    t = self.get_target_tensor(d, target_is_real)
                    l = -torch.mean(torch.min(d * t - 1, torch.zeros_like(d))) if target_is_real else \
                        -torch.mean(torch.min(-d - 1, torch.zeros_like(d)))
                else:
                    l = -torch.mean(d)
            else:
                l = self.loss_fct(d.float(), self.get_target_tensor(d, target_is_real).float())
            losses.append(l)
        if self.reduction == "mean":
            return torch.mean(torch.stack(losses))
        if self.reduction == "sum":
            return torch.sum(torch.stack(losses))
        return losses[0] if len(losses) == 1 else losses
