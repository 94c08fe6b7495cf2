"""generative.inferers (train_ldm.py:28, train_ddpm.py:17): the B200 inferers."""
from medical_image_generation_b200.inferers import DiffusionInferer, LatentDiffusionInferer  # noqa: F401

__all__ = ["DiffusionInferer", "LatentDiffusionInferer"]
