"""generative.losses.perceptual.medicalnet_intensity_normalisation (train_ldm.py:31; only used by the 2-D FID validation)."""


def medicalnet_intensity_normalisation(volume):
    """Per-volume z-score, as MedicalNet expects (upstream one-liner)."""
    return (volume - volume.mean()) / volume.std()
