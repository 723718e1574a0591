"""Minimal stand-in for `tensorly` (absent from this image, no network).

TEST INFRASTRUCTURE ONLY. Provides exactly the four symbols the unmodified reference file
`neuralop/layers/spectral_convolution.py:8-17,22,46` touches: `set_backend`, `ndim`, `einsum`
and the `plugins.use_opt_einsum` hook. `tl.einsum` on the pytorch backend is `torch.einsum`.
"""
import torch

# the opt_einsum stand-in next to this package must not be picked up by torch.einsum's path optimiser
torch.backends.opt_einsum.enabled = False

from . import plugins  # noqa: F401


def set_backend(name):
    return None


def ndim(x):
    return x.ndim


def einsum(eq, *operands):
    return torch.einsum(eq, *operands)
