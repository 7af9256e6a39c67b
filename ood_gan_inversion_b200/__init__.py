"""ood_gan_inversion_b200 -- B200 (sm_100a) hot path of OOD-GAN-inversion behind the reference's module API.

CUDA-only: importing works anywhere, but every op raises if libood_b200.so is missing or a tensor is not on a CUDA
device (there is no CPU / PyTorch fallback).
"""
from . import _lib, kernels  # noqa: F401

__all__ = ['kernels']
