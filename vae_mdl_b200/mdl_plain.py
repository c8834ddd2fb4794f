"""``PixelMixtureDiscretizedLogistic`` -- host-side mirror of utils/mdl_plain.py:7-121 (+ ``get_mixture_params`` :124-168).

The pixel mixture of discretized logistics that does NOT condition on the observed x: the green / blue means are
chained on the component's own red / green means.  Same kernels as ``MixtureDiscretizedLogistic`` (instantiated with
the other chain), same parameter row ``[logit(M) | (mu, s, kappa) x RGB]``.
"""
from __future__ import annotations

import math

import torch

from . import _abi
from . import functional as F
from ._noise import uniform_noise

__all__ = ["PixelMixtureDiscretizedLogistic", "get_mixture_params"]


def get_mixture_params(parameters: torch.Tensor):
    """utils/mdl_plain.py:124-168 -> ``loc, logscale [..., 3, n_mix]``, ``mix_logits [..., n_mix]``.  A host-side view
    helper for callers that want the tensors (diagnostics, plots); the kernels never materialise them."""
    _abi.require_cuda(parameters, "parameters")
    shape = list(parameters.shape)
    n_mix = shape[-1] // 10
    mix_logits = parameters[..., :n_mix]                                                  # :143
    rest = parameters[..., n_mix:].reshape(shape[:-1] + [3, 3 * n_mix])                   # :146
    _loc, logscale, coeffs = torch.split(rest, n_mix, dim=-1)                             # :149
    logscale = torch.clamp(logscale, min=-7.0)                                            # :150
    coeffs = torch.tanh(coeffs)                                                           # :151
    loc_r = _loc[..., 0, :]                                                               # :160
    loc_g = _loc[..., 1, :] + coeffs[..., 0, :] * loc_r                                   # :161
    loc_b = _loc[..., 2, :] + coeffs[..., 1, :] * loc_r + coeffs[..., 2, :] * loc_g       # :162
    return torch.stack([loc_r, loc_g, loc_b], dim=-2), logscale, mix_logits              # :164-168


class PixelMixtureDiscretizedLogistic:
    def __init__(self, parameters: torch.Tensor, low=-1.0, high=1.0, levels=256.0):
        """``parameters [..., batch, h, w, n_mix * 10]`` (utils/mdl_plain.py:18-34)."""
        _abi.require_cuda(parameters, "parameters")
        if not float(levels) > 1.0 or not float(high) > float(low):
            raise ValueError("need high > low and levels > 1")
        if (float(low), float(high), float(levels)) != (-1.0, 1.0, 256.0) and \
                (float(high) - float(low)) / (float(levels) - 1.0) > 2.0 * 27.0 / math.exp(7.0):
            raise ValueError("bin width (high - low) / (levels - 1) must not exceed 0.049 (e.g. levels >= 42 on [-1, 1]): "
                             "coarser grids are outside the float32 range of the kernels' linear-domain evaluation")
        self._parameters = parameters
        self._bins = (float(low), float(high), float(levels))
        self.n_mix = parameters.shape[-1] // 10                                           # :34
        if self.n_mix * 10 != parameters.shape[-1] or self.n_mix < 1:
            raise ValueError(f"last dim must be n_mix * 10, got {parameters.shape[-1]}")
        self.low, self.high, self.levels = low, high, levels
        self.interval_width = (high - low) / (levels - 1.0)                               # discretized_logistic.py:18
        self.dx = self.interval_width / 2.0                                               # :21
        self.axes = [-1, -2]          # log_prob returns [..., h, w]: the caller sums over the two image axes

    # lazily materialised views of the reference attributes (utils/mdl_plain.py:27-33)
    @property
    def loc(self):
        return get_mixture_params(self._parameters)[0]

    @property
    def logscale(self):
        return get_mixture_params(self._parameters)[1]

    @property
    def mix_logits(self):
        return self._parameters[..., :self.n_mix]

    def log_prob(self, x: torch.Tensor) -> torch.Tensor:
        """x in [0,1] ``[batch, h, w, 3]``; returns ``[..., h, w]`` (utils/mdl_plain.py:36-66, no trailing 1)."""
        return F.modl_log_prob(self._parameters, x, plain=self._bins)

    def log_likelihood(self, x: torch.Tensor, dtype: torch.dtype = torch.float32) -> torch.Tensor:
        """Sum of ``log_prob`` over the image, fused into the kernel (per-pixel tensor never written)."""
        return F.modl_log_likelihood(self._parameters, x, dtype=dtype, plain=self._bins)

    def sample(self, n_samples=[], u_mix=None, u_log=None, generator=None, return_index=False, return_quantised=False):
        """utils/mdl_plain.py:68-102: ``n_samples=[]`` -> ``[..., h, w, 3]``, ``n`` / ``[n]`` -> leading ``[n]``; values in
        [0,1].  ``u_mix [n, ..., h, w, n_mix]`` selects the component, ``u_log [n, ..., h, w, 3, n_mix]`` drives the
        logistic draw of every component (:86-88); drawn on the device when omitted."""
        p = self._parameters
        lead = tuple(p.shape[:-1])
        if isinstance(n_samples, (list, tuple)):
            ns = tuple(int(v) for v in n_samples)
        else:
            ns = (int(n_samples),)
        if len(ns) > 1:
            raise ValueError("only scalar sample shapes are supported")
        n = ns[0] if ns else 1
        if u_mix is None:
            u_mix = uniform_noise((n,) + lead + (self.n_mix,), p.device, generator)
        if u_log is None:
            u_log = uniform_noise((n,) + lead + (3, self.n_mix), p.device, generator)
        out = F.modl_sample(p, u_mix.reshape((n,) + lead + (self.n_mix,)), u_log.reshape((n,) + lead + (3, self.n_mix)),
                            _abi.SAMPLE_PLAIN, _abi.RANGE_UNIT, want_quantised=return_quantised, want_index=return_index,
                            clip=self._bins[:2])
        outs = out if isinstance(out, tuple) else (out,)
        if not ns:
            outs = tuple(o[0] for o in outs)
        return outs if len(outs) > 1 else outs[0]

    def mean(self, u_mix=None, generator=None, **kwargs):
        """utils/mdl_plain.py:104-121: the (clipped) locations of ONE categorically sampled component per pixel, in [0,1]."""
        p = self._parameters
        lead = tuple(p.shape[:-1])
        if u_mix is None:
            u_mix = uniform_noise((1,) + lead + (self.n_mix,), p.device, generator)
        return F.modl_sample(p, u_mix.reshape((1,) + lead + (self.n_mix,)), None, _abi.SAMPLE_PLAIN, _abi.RANGE_UNIT)[0]
