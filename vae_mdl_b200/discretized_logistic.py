"""``DiscretizedLogistic`` -- mirror of utils/discretized_logistic.py:5-88 (plain class, not a distribution subclass)."""
from __future__ import annotations

import torch

from . import _abi
from . import functional as F
from ._noise import uniform_noise

__all__ = ["DiscretizedLogistic"]


class DiscretizedLogistic:
    """Discretized logistic ``f(x; mu, s)``; ``loc``/``logscale`` may be the two halves of one ``[..., 6]`` tensor
    (``torch.split`` / ``chunk`` views, models/model03.py:88-91) -- they are then read in place, without a copy."""

    def __init__(self, loc, logscale, low=-1.0, high=1.0, levels=256.0):
        _abi.require_cuda(loc, "loc")
        self.loc = loc                                        # :11
        self.logscale = logscale                              # :12
        self.low = low
        self.high = high
        self.levels = levels
        self.interval_width = (high - low) / (levels - 1.0)   # :18
        self.dx = self.interval_width / 2.0                   # :21
        self.axes = [-1, -2, -3]                              # assigned by the models (models/model03.py:143)

    def _iwae_spec(self):
        return "dl", {"low": self.low, "high": self.high, "levels": self.levels}, self.loc, self.logscale

    def logistic_cdf(self, x):
        """:23-25 (diagnostic helper, not on the hot path): sigmoid((x-loc) exp(-logscale))."""
        return torch.sigmoid((x - self.loc) * torch.exp(-self.logscale))

    def logistic_log_prob_approx(self, x):
        """:27-33 (diagnostic helper; the kernels evaluate this branch internally)."""
        a = (x - self.loc) / torch.exp(self.logscale)
        import math
        return -a - self.logscale - 2 * torch.nn.functional.softplus(-a) + math.log(self.interval_width)

    def log_prob(self, x):
        """Element-wise, shape of ``loc`` (:35-78); x is used as given (no rescale)."""
        return F.dlogistic_log_prob(self.loc, self.logscale, x, self.low, self.high, self.levels)

    def log_likelihood(self, x, n_event_dims=3, dtype=torch.float32):
        """``reduce_sum(log_prob(x), last n_event_dims axes)`` without the element-wise tensor (models/loss.py:32)."""
        return F.dlogistic_log_likelihood(self.loc, self.logscale, x, self.low, self.high, self.levels, n_event_dims, dtype)

    def sample(self, n_samples=[], u=None, generator=None):
        """:80-85; ``n_samples=[]`` -> shape of ``loc``; ``n`` or ``[n]`` -> leading ``[n]``."""
        if isinstance(n_samples, (list, tuple)):
            lead = tuple(int(v) for v in n_samples)
        else:
            lead = (int(n_samples),)
        if u is None:
            u = uniform_noise(lead + tuple(self.loc.shape), self.loc.device, generator)
        return F.dlogistic_sample(self.loc, self.logscale, u, self.low, self.high).reshape(lead + tuple(self.loc.shape))

    def mean(self, **kwargs):
        """:87-88."""
        return self.loc
