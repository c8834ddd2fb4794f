"""``MixtureDiscretizedLogisticOpenaiIWAE`` -- mirror of utils/mdl_openai_iwae.py:16-102.

The reference folds the leading importance-sample dims into the batch, tiles x ``S`` times and maps x from [0,1] to
[-1,1] (:35-56).  Here the fold is index arithmetic inside the kernel (image ``n = s*B + b`` reads ``x[n % B]``).
"""
from __future__ import annotations

import torch

from . import _abi
from . import functional as F
from ._noise import sample_shape_to_n, uniform_noise
from .mdl_openai import (discretized_mix_logistic_loss, int_shape, log_prob_from_logits, log_sum_exp,  # noqa: F401
                         sample_from_discretized_mix_logistic)

__all__ = ["MixtureDiscretizedLogisticOpenaiIWAE"]


class MixtureDiscretizedLogisticOpenaiIWAE:
    def __init__(self, logits: torch.Tensor):
        _abi.require_cuda(logits, "logits")
        self.logits = logits                                  # :29
        self.shape = list(logits.shape)                       # :30
        self.n_mix = self.shape[-1] // 10                     # :31
        self.dtype = logits.dtype
        self._axes = [-1, -2, -3]

    def _iwae_spec(self):
        return "modl", {"x_range": _abi.RANGE_UNIT, "edge_mode": _abi.EDGE_OPENAI, "plain": False}, self.logits, None

    def _check_x(self, x):
        # the reference computes repeats = (S*B) // x.shape[0] (:49): x MUST carry its batch dim
        if x.dim() != 4:
            raise ValueError("MixtureDiscretizedLogisticOpenaiIWAE.log_prob needs x with its batch dim [B,H,W,3] "
                             "(utils/mdl_openai_iwae.py:49)")

    def log_prob(self, x):
        """x in [0,1] ``[B,H,W,3]`` -> ``[S..., B, H, W, 1]`` (:33-67)."""
        self._check_x(x)
        return F.modl_log_prob(self.logits, x, _abi.RANGE_UNIT, _abi.EDGE_OPENAI).unsqueeze(-1)

    def log_likelihood(self, x, dtype=torch.float32):
        self._check_x(x)
        return F.modl_log_likelihood(self.logits, x, _abi.RANGE_UNIT, _abi.EDGE_OPENAI, dtype)

    def sample(self, sample_shape=(), u_mix=None, u_log=None, generator=None, **kw):
        """-> ``[n, S..., B, H, W, 3]`` in [0,1] (:69-99)."""
        n, squeeze = sample_shape_to_n(sample_shape)
        l = self.logits
        lead = tuple(l.shape[:-1])
        if u_mix is None:
            u_mix = uniform_noise((n,) + lead + (self.n_mix,), l.device, generator)
        if u_log is None:
            u_log = uniform_noise((n,) + lead + (3,), l.device, generator)
        u_mix = u_mix.reshape((n,) + lead + (self.n_mix,))
        u_log = u_log.reshape((n,) + lead + (3,))
        out = F.modl_sample(l, u_mix, u_log, _abi.SAMPLE_OPENAI, _abi.RANGE_UNIT,
                            want_quantised=kw.get("return_quantised", False), want_index=kw.get("return_index", False))
        outs = out if isinstance(out, tuple) else (out,)
        if squeeze:
            outs = tuple(o[0] for o in outs)
        return outs if len(outs) > 1 else outs[0]

    def _sample_n(self, n, seed=None, **kwargs):
        return self.sample(n, **kwargs)

    def mean(self, n=100, **kwargs):
        """:101-102."""
        return self.sample(n, **kwargs).mean(dim=0)

    @property
    def axes(self):
        return self._axes

    @axes.setter
    def axes(self, axes):
        self._axes = axes
