"""generative.losses (train_autoencoder.py:27): both need downloaded networks -> placeholders."""
from .._placeholder import placeholder

PatchAdversarialLoss = placeholder("losses.PatchAdversarialLoss", "adversarial AE loss, SURVEY.md 8f-3")
PerceptualLoss = placeholder("losses.PerceptualLoss", "needs downloaded LPIPS / MedicalNet weights")
