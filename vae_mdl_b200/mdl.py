"""``MixtureDiscretizedLogistic`` -- host-side mirror of the reference class (utils/mdl.py:19-263).

Same constructor, methods, attribute names and shape semantics; tensors are ``torch`` CUDA tensors and all arithmetic
runs in the sm_100a kernels of ``libvaemdl_b200.so`` (no TensorFlow, no PyTorch math, no CPU path).
"""
from __future__ import annotations

import torch

from . import _abi
from . import functional as F
from ._noise import sample_shape_to_n, uniform_noise

__all__ = ["MixtureDiscretizedLogistic"]


class MixtureDiscretizedLogistic:
    def __init__(self, parameters: torch.Tensor):
        """Assumes parameters shape ``[?] + [batch, h, w, n_mix * 10]`` (utils/mdl.py:20-54)."""
        _abi.require_cuda(parameters, "parameters")
        self._parameters = parameters
        self.shape = list(parameters.shape)
        self.n_mix = self.shape[-1] // 10                     # utils/mdl.py:44
        if self.n_mix * 10 != self.shape[-1] or self.n_mix < 1:
            raise ValueError(f"last dim must be n_mix * 10, got {self.shape[-1]}")
        self.interval_width = 2.0 / 255.0                     # :47
        self.dx = self.interval_width / 2.0                   # :50
        self.low, self.high = -1.0, 1.0                       # :52
        self._axes = [-1, -2, -3]                             # :54
        self.dtype = parameters.dtype

    def _iwae_spec(self):
        """What ``loss.iwae_loss`` needs to run the fused objective on this observation model."""
        return "modl", {"x_range": _abi.RANGE_UNIT, "edge_mode": _abi.EDGE_MDL, "plain": False}, self._parameters, None

    # ---- densities -------------------------------------------------------------------------------------------
    def log_prob(self, x: torch.Tensor) -> torch.Tensor:
        """x in [0,1], ``[batch,h,w,3]`` or broadcastable ``[h,w,3]``; returns ``[..., h, w, 1]`` (utils/mdl.py:56-92)."""
        return F.modl_log_prob(self._parameters, x, _abi.RANGE_UNIT, _abi.EDGE_MDL).unsqueeze(-1)

    def log_likelihood(self, x: torch.Tensor, dtype: torch.dtype = torch.float32) -> torch.Tensor:
        """``reduce_sum(log_prob(x), [-1,-2,-3])`` (models/loss.py:32) in one kernel, per-pixel tensor never written."""
        return F.modl_log_likelihood(self._parameters, x, _abi.RANGE_UNIT, _abi.EDGE_MDL, dtype)

    # ---- sampling ----------------------------------------------------------------------------------------------
    def sample(self, sample_shape=(), u_mix=None, u_log=None, generator=None, return_index=False, return_quantised=False):
        """tfd semantics: ``sample()`` -> parameters.shape[:-1] + [3]; ``sample(n)`` -> ``[n, ...]`` (utils/mdl.py:209-252).

        ``u_mix [n, ..., h, w, n_mix]`` / ``u_log [n, ..., h, w, 3, n_mix]``: explicit uniforms (a logistic draw for every
        mixture, :213); drawn on the device in (1e-5, 1-1e-5) when omitted.  Output in [0,1] (:250).
        """
        n, squeeze = sample_shape_to_n(sample_shape)
        p = self._parameters
        lead = tuple(p.shape[:-1])
        if u_mix is None:
            u_mix = uniform_noise((n,) + lead + (self.n_mix,), p.device, generator)
        if u_log is None:
            u_log = uniform_noise((n,) + lead + (3, self.n_mix), p.device, generator)
        u_mix = u_mix.reshape((n,) + lead + (self.n_mix,))
        u_log = u_log.reshape((n,) + lead + (3, self.n_mix))
        out = F.modl_sample(p, u_mix, u_log, _abi.SAMPLE_MDL, _abi.RANGE_UNIT,  # parameters are NOT tiled n times
                            want_quantised=return_quantised, want_index=return_index)
        outs = out if isinstance(out, tuple) else (out,)
        if squeeze:
            outs = tuple(o[0] for o in outs)
        return outs if len(outs) > 1 else outs[0]

    def _sample_n(self, n, seed=None, **kwargs):
        return self.sample(n, **kwargs)

    def mean(self, n=100, **kwargs):
        """Monte-Carlo mean of ``n`` samples (utils/mdl.py:254-255)."""
        return self.sample(n, **kwargs).mean(dim=0)

    # ---- reduction axes used by the loss (utils/mdl.py:257-263) ---------------------------------------------------
    @property
    def axes(self):
        return self._axes

    @axes.setter
    def axes(self, axes):
        self._axes = axes

    @property
    def parameters(self):
        return self._parameters
