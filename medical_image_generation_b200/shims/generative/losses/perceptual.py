"""generative.losses.perceptual (train_autoencoder.py:27,416,601; train_ldm.py:31).

`PerceptualLoss(spatial_dims, network_type="vgg", is_fake_3d=True, fake_3d_ratio=...)` as the reference configures it
(configuration.py:961-964: 2-D `vgg`; 3-D `vgg` with is_fake_3d and fake_3d_ratio 0.2): LPIPS on a VGG-16 trunk, applied
to a random subset of the slices along each of the three axes for volumes. Restated from the published definitions of
monai-generative's `PerceptualLoss` and of the `lpips` package [upstream-memory; PARITY UNPINNED -- neither package is
installed here nor under /root/reference; the VGG-16 trunk's layer layout IS pinned, to torchvision's definition:
tests/test_host_logic.py::test_lpips_vgg_trunk_matches_torchvision_vgg16_layout]. Outside the north star's hot path: plain torch.nn (ATen / cuDNN), like the
discriminator.

Weights: there is no network in this image, so nothing is downloaded. `pretrained=True` (the default, as upstream) needs
a local file -- `pretrained_path=` or $MEDIMGEN_LPIPS_WEIGHTS -- holding the state dict of `lpips.LPIPS(net='vgg')`
(keys `net.slice*.N.{weight,bias}`, `lin*.model.1.weight`; produce it once on a connected machine with
`torch.save(lpips.LPIPS(net='vgg').state_dict(), path)`); without one the constructor raises. `pretrained=False` gives a
randomly initialised network (tests, smoke runs)."""
from __future__ import annotations

import os

import torch
from torch import nn


def medicalnet_intensity_normalisation(volume):
    """Per-volume z-score, as MedicalNet expects (upstream one-liner)."""
    return (volume - volume.mean()) / volume.std()


# torchvision vgg16().features indices of the conv layers inside each LPIPS slice (relu1_2, relu2_2, relu3_3, relu4_3,
# relu5_3); a max-pool opens slices 2-5
_VGG_SLICES = (((0, 3, 64), (2, 64, 64)),
               ((5, 64, 128), (7, 128, 128)),
               ((10, 128, 256), (12, 256, 256), (14, 256, 256)),
               ((17, 256, 512), (19, 512, 512), (21, 512, 512)),
               ((24, 512, 512), (26, 512, 512), (28, 512, 512)))
_CHANNELS = (64, 128, 256, 512, 512)


class _ScalingLayer(nn.Module):
    def __init__(self):
        super().__init__()
        self.register_buffer("shift", torch.tensor([-0.030, -0.088, -0.188])[None, :, None, None])
        self.register_buffer("scale", torch.tensor([0.458, 0.448, 0.450])[None, :, None, None])

    def forward(self, x):
        return (x - self.shift) / self.scale      # a 1-channel image broadcasts to the 3 channels the trunk expects


class _Vgg16Trunk(nn.Module):
    def __init__(self):
        super().__init__()
        for k, convs in enumerate(_VGG_SLICES, start=1):
            seq = nn.Sequential()
            if k > 1:
                seq.add_module(str(convs[0][0] - 1), nn.MaxPool2d(2, 2))
            for idx, cin, cout in convs:
                seq.add_module(str(idx), nn.Conv2d(cin, cout, 3, padding=1))
                seq.add_module(str(idx + 1), nn.ReLU(inplace=False))
            setattr(self, f"slice{k}", seq)

    def forward(self, x):
        feats = []
        for k in range(1, 6):
            x = getattr(self, f"slice{k}")(x)
            feats.append(x)
        return feats


class _NetLinLayer(nn.Module):
    def __init__(self, chn_in: int, use_dropout: bool = True):
        super().__init__()
        layers = [nn.Dropout()] if use_dropout else []
        layers.append(nn.Conv2d(chn_in, 1, 1, bias=False))
        if not use_dropout:
            layers.insert(0, nn.Identity())      # keep the conv at index 1 (`lin*.model.1.weight`)
        self.model = nn.Sequential(*layers)

    def forward(self, x):
        return self.model(x)


class LPIPS(nn.Module):
    """Learned perceptual image patch similarity, VGG-16 variant: d(x, y) = sum_l mean_hw( w_l . (f_l(x)^ - f_l(y)^)^2 )
    with channel-unit-normalised features f^ and non-negative 1x1 weights w_l. Returns (N, 1, 1, 1)."""

    def __init__(self, pretrained: bool = True, net: str = "vgg", pretrained_path: str | None = None,
                 state_dict_key: str | None = None):
        super().__init__()
        if net != "vgg":
            raise NotImplementedError(f"LPIPS trunk {net!r}: the reference configures 'vgg' (configuration.py:961-964)")
        self.scaling_layer = _ScalingLayer()
        self.net = _Vgg16Trunk()
        for k, c in enumerate(_CHANNELS):
            setattr(self, f"lin{k}", _NetLinLayer(c))
        if pretrained:
            path = pretrained_path or os.environ.get("MEDIMGEN_LPIPS_WEIGHTS")
            if not path or not os.path.isfile(path):
                raise RuntimeError(
                    "PerceptualLoss(pretrained=True) needs the LPIPS-VGG weights as a local file (no network here): pass "
                    "pretrained_path= or set MEDIMGEN_LPIPS_WEIGHTS to the state dict of lpips.LPIPS(net='vgg'); "
                    "pretrained=False builds a randomly initialised network")
            sd = torch.load(path, map_location="cpu")
            if state_dict_key is not None:
                sd = sd[state_dict_key]
            sd = {k: v for k, v in sd.items() if not k.startswith("lins.")}     # lpips keeps the lin layers twice
            missing, unexpected = self.load_state_dict(sd, strict=False)
            missing = [k for k in missing if not k.startswith("scaling_layer.")]
            if missing or unexpected:
                raise RuntimeError(f"LPIPS weights at {path} do not match: missing {missing[:4]}, unexpected {unexpected[:4]}")
        else:
            with torch.no_grad():
                for k in range(len(_CHANNELS)):      # the published weights are non-negative; keep that property
                    getattr(self, f"lin{k}").model[1].weight.abs_()
        for p in self.parameters():
            p.requires_grad_(False)
        self.eval()

    def train(self, mode: bool = True):
        return super().train(False)      # a frozen metric network: dropout stays off

    @staticmethod
    def _unit(x, eps: float = 1e-10):
        return x / (torch.sqrt(torch.sum(x * x, dim=1, keepdim=True)) + eps)

    def forward(self, in0, in1, normalize: bool = False):
        if normalize:
            in0, in1 = 2 * in0 - 1, 2 * in1 - 1
        f0, f1 = self.net(self.scaling_layer(in0)), self.net(self.scaling_layer(in1))
        total = 0
        for k, (a, b) in enumerate(zip(f0, f1)):
            diff = (self._unit(a) - self._unit(b)) ** 2
            total = total + getattr(self, f"lin{k}")(diff).mean([2, 3], keepdim=True)
        return total


class PerceptualLoss(nn.Module):
    def __init__(self, spatial_dims: int, network_type: str = "alex", is_fake_3d: bool = True, fake_3d_ratio: float = 0.5,
                 cache_dir: str | None = None, pretrained: bool = True, pretrained_path: str | None = None,
                 pretrained_state_dict_key: str | None = None):
        super().__init__()
        if spatial_dims not in (2, 3):
            raise NotImplementedError("Perceptual loss is implemented only in 2D and 3D.")
        if spatial_dims == 3 and not is_fake_3d and "medicalnet_" not in network_type:
            raise ValueError("MedicalNet networks are only compatible with ``spatial_dims=3``."
                             "Argument is_fake_3d must be set to False.")
        if network_type != "vgg":
            raise NotImplementedError(f"network_type {network_type!r} needs monai-generative; this package ships the "
                                      "LPIPS-VGG variant the reference configures (configuration.py:961-964)")
        self.spatial_dims, self.is_fake_3d, self.fake_3d_ratio = spatial_dims, is_fake_3d, fake_3d_ratio
        self.perceptual_function = LPIPS(pretrained=pretrained, net="vgg", pretrained_path=pretrained_path,
                                         state_dict_key=pretrained_state_dict_key)

    def _calculate_axis_loss(self, input, target, spatial_axis: int):
        """The loss over a random `fake_3d_ratio` share of the 2-D slices perpendicular to `spatial_axis`."""
        def batchify_axis(x, perm):
            s = x.float().permute((0,) + perm).contiguous()
            return s.view(-1, x.shape[perm[1]], x.shape[perm[2]], x.shape[perm[3]])

        kept = [2, 3, 4]
        kept.remove(spatial_axis)
        perm = (spatial_axis, 1) + tuple(kept)
        in_s, tg_s = batchify_axis(input, perm), batchify_axis(target, perm)
        idx = torch.randperm(in_s.shape[0])[: int(in_s.shape[0] * self.fake_3d_ratio)].to(in_s.device)
        return torch.mean(self.perceptual_function(torch.index_select(in_s, 0, idx), torch.index_select(tg_s, 0, idx)))

    def forward(self, input: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        if target.shape != input.shape:
            raise ValueError(f"ground truth has differing shape ({target.shape}) from input ({input.shape})")
        if self.spatial_dims == 3 and self.is_fake_3d:
            loss = sum(self._calculate_axis_loss(input, target, spatial_axis=a) for a in (2, 3, 4))
        else:
            loss = self.perceptual_function(input.float(), target.float())
        return torch.mean(loss)
