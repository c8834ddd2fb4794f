"""``MixtureDiscretizedLogisticOpenai`` and the PixelCNN++ module functions -- mirror of utils/mdl_openai.py.

``discretized_mix_logistic_loss`` / ``sample_from_discretized_mix_logistic`` keep the reference signatures
(utils/mdl_openai.py:83, :160); the work is done by the fused sm_100a kernels.
"""
from __future__ import annotations

import math

import torch

from . import _abi
from . import functional as F
from ._noise import sample_shape_to_n, uniform_noise

__all__ = [
    "MixtureDiscretizedLogisticOpenai",
    "discretized_mix_logistic_loss",
    "sample_from_discretized_mix_logistic",
    "log_sum_exp",
    "log_prob_from_logits",
    "int_shape",
]


def int_shape(x):
    """utils/mdl_openai.py:64-65."""
    return list(map(int, x.shape))


def log_sum_exp(x: torch.Tensor) -> torch.Tensor:
    """Stable log-sum-exp over the last axis (utils/mdl_openai.py:68-73) = logmeanexp + log(n), on the lme kernel."""
    return F.logmeanexp(x, -1) + math.log(x.shape[-1])


def log_prob_from_logits(x: torch.Tensor) -> torch.Tensor:
    """Stable log-softmax over the last axis (utils/mdl_openai.py:76-80)."""
    return x - log_sum_exp(x).unsqueeze(-1)


def discretized_mix_logistic_loss(x: torch.Tensor, l: torch.Tensor, sum_all: bool = True) -> torch.Tensor:
    """x in [-1,1] ``[N,H,W,3]``, l ``[N,H,W,10*nr_mix]`` (utils/mdl_openai.py:83-157).
    ``sum_all=True`` returns the NEGATIVE total log-likelihood (:154); ``False`` the per-pixel log-prob ``[N,H,W]`` (:157)."""
    if l.dim() != 4 or x.dim() != 4:
        raise ValueError("discretized_mix_logistic_loss expects 4-D x and l (utils/mdl_openai.py:85-86)")
    if sum_all:
        return -F.modl_log_likelihood(l, x, _abi.RANGE_SYM, _abi.EDGE_OPENAI).sum()
    return F.modl_log_prob(l, x, _abi.RANGE_SYM, _abi.EDGE_OPENAI)


def sample_from_discretized_mix_logistic(l: torch.Tensor, nr_mix: int, u_mix=None, u_log=None, generator=None,
                                         return_index=False, return_quantised=False):
    """``l [N,H,W,10*nr_mix]`` -> ``[N,H,W,3]`` in [-1,1] (utils/mdl_openai.py:160-193).  Explicit-noise form of the
    commented PixelCNN++ lines (:167, :185-186): ``u_mix [N,H,W,nr_mix]``, ``u_log [N,H,W,3]``."""
    if l.shape[-1] != 10 * nr_mix:
        raise ValueError("last dim of l must be 10 * nr_mix")
    lead = tuple(l.shape[:-1])
    if u_mix is None:
        u_mix = uniform_noise(lead + (nr_mix,), l.device, generator)
    if u_log is None:
        u_log = uniform_noise(lead + (3,), l.device, generator)
    return F.modl_sample(l, u_mix, u_log, _abi.SAMPLE_OPENAI, _abi.RANGE_SYM, want_quantised=return_quantised,
                         want_index=return_index)


class MixtureDiscretizedLogisticOpenai:
    """utils/mdl_openai.py:15-58: 4-D logits, x in [-1,1], samples in [-1,1]."""

    def __init__(self, logits: torch.Tensor):
        _abi.require_cuda(logits, "logits")
        if logits.dim() != 4:
            raise ValueError("MixtureDiscretizedLogisticOpenai expects logits [batch, h, w, n_mix * 10]")
        self.logits = logits                                  # :28
        self.n_mix = logits.shape[-1] // 10                   # :29
        self.dtype = logits.dtype
        self._axes = [-1, -2]

    def log_prob(self, x):
        """-> ``[B,H,W]`` (utils/mdl_openai.py:31-32)."""
        return discretized_mix_logistic_loss(x, self.logits, sum_all=False)

    def sample(self, sample_shape=(), u_mix=None, u_log=None, generator=None, **kw):
        """-> ``[n,B,H,W,3]`` in [-1,1]; the n-fold tiling of the logits (:39-45) is an index computation here."""
        n, squeeze = sample_shape_to_n(sample_shape)
        l = self.logits
        lead = tuple(l.shape[:-1])
        u_mix = uniform_noise((n,) + lead + (self.n_mix,), l.device, generator) if u_mix is None else \
            u_mix.reshape((n,) + lead + (self.n_mix,))
        u_log = uniform_noise((n,) + lead + (3,), l.device, generator) if u_log is None else \
            u_log.reshape((n,) + lead + (3,))
        out = F.modl_sample(l, u_mix, u_log, _abi.SAMPLE_OPENAI, _abi.RANGE_SYM,
                            want_quantised=kw.get("return_quantised", False), want_index=kw.get("return_index", False))
        outs = out if isinstance(out, tuple) else (out,)                                          # :53
        if squeeze:
            outs = tuple(o[0] for o in outs)
        return outs if len(outs) > 1 else outs[0]

    def _sample_n(self, n, seed=None, **kwargs):
        return self.sample(n, **kwargs)

    def mean(self, n=100, **kwargs):
        """utils/mdl_openai.py:57-58."""
        return self.sample(n, **kwargs).mean(dim=0)

    @property
    def axes(self):
        return self._axes

    @axes.setter
    def axes(self, axes):
        self._axes = axes
