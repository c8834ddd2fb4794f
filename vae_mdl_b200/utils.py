"""Mirror of the hot-path part of utils/utils.py: ``logmeanexp`` (:9-11) and the distribution containers (:48-96)."""
from __future__ import annotations

from collections import namedtuple
from typing import Any, NamedTuple, Optional

import torch

from . import functional as F

__all__ = ["logmeanexp", "Dist", "DistributionTuple"]


def logmeanexp(log_w: torch.Tensor, axis: int) -> torch.Tensor:
    """``log(mean(exp(log_w - max), axis)) + max`` (utils/utils.py:9-11) on the sm_100a kernel; differentiable
    (gradient = softmax over ``axis``, as the reference's un-stopped max gives)."""
    return F.logmeanexp(log_w, axis)


class Dist(namedtuple("Dist", "dist sample axes")):
    """utils/utils.py:48-71."""

    @property
    def z(self):
        return self.sample

    @property
    def x(self):
        return self.sample

    @property
    def p(self):
        return self.dist

    @property
    def q(self):
        return self.dist


class DistributionTuple(NamedTuple):
    """utils/utils.py:83-96."""

    dist: Any
    sample: Optional[torch.Tensor] = None
    axes: tuple = (-1, -2, -3)

    @property
    def z(self):
        return self.sample

    @property
    def x(self):
        return self.sample
