"""generative.metrics (train_ldm.py:32): FID / MMD / SSIM validation metrics, 2-D only and fed by torch.hub networks
(train_ldm.py:241-330, 547) -- outside the hot path."""
from ._placeholder import placeholder

FIDMetric = placeholder("metrics.FIDMetric", "validation metric, needs torch.hub feature networks")
MMDMetric = placeholder("metrics.MMDMetric", "validation metric")
SSIMMetric = placeholder("metrics.SSIMMetric", "validation metric")
MultiScaleSSIMMetric = placeholder("metrics.MultiScaleSSIMMetric", "validation metric")
