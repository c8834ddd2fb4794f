"""Uniform noise for the explicit-noise samplers, and tfd ``sample_shape`` handling."""
from __future__ import annotations

import torch


def uniform_noise(shape, device, generator=None) -> torch.Tensor:
    """U(1e-5, 1 - 1e-5), the range of the original PixelCNN++ sampler (utils/mdl_openai.py:167, :185)."""
    u = torch.rand(shape, device=device, dtype=torch.float32, generator=generator)
    return u.mul_(1.0 - 2e-5).add_(1e-5)


def sample_shape_to_n(sample_shape):
    """tfd: ``sample(())`` calls ``_sample_n(1)`` and drops the leading dim; ``sample(n)`` / ``sample([n])`` keep it."""
    if isinstance(sample_shape, (tuple, list)):
        if len(sample_shape) == 0:
            return 1, True
        if len(sample_shape) == 1:
            return int(sample_shape[0]), False
        raise ValueError("only scalar or 1-D sample shapes are supported")
    return int(sample_shape), False
