"""generative.networks.nets: VQVAE (train_ldm.py:30, optional `-l vq` latent space) and PatchDiscriminator
(train_autoencoder.py:26, adversarial loss) are outside the hot path; DiffusionModelUNet (train_ddpm.py:18 builds the
STOCK MONAI U-Net) maps to the strided B200 U-Net, which reduces to the stock architecture for isotropic stride-2 levels."""
from .._placeholder import placeholder

VQVAE = placeholder("networks.nets.VQVAE", "the VQ-VAE latent space is not part of BASELINE.json's north star")
PatchDiscriminator = placeholder("networks.nets.PatchDiscriminator", "adversarial AE loss, SURVEY.md 8f-3")
