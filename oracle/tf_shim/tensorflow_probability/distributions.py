"""`tfd.Distribution` dispatch + the four stock distributions the reference's path instantiates.

TEST INFRASTRUCTURE ONLY.  Semantics follow TFP v0.16 (the version the reference links, utils/mdl_openai.py:35):
* `Distribution.log_prob(x)` -> `_log_prob(convert_to_tensor(x))`; `mean(**kw)` -> `_mean(**kw)`;
  `sample(sample_shape=())` -> `_sample_n(prod(sample_shape))` reshaped to `sample_shape + batch/event shape`
  (distribution.py `_call_sample_n`: a scalar `()` request runs `_sample_n(1)` and drops the leading 1);
* `Logistic._sample_n`: `u ~ U(tiny, 1)`, `loc + scale * (log u - log1p(-u))` (logistic.py:160);
* `OneHotCategorical` / `Categorical` sampling is Gumbel-max over the logits (what `tf.random.categorical`'s kernel
  does and what the commented PixelCNN++ line utils/mdl_openai.py:167 spells out); the first maximum wins ties.
Uniform noise never comes from an RNG here: the caller queues it with `push_uniforms` in the order the reference's
code will draw it, so the reference's samplers become deterministic functions that fixtures can pin.
"""
from __future__ import annotations

import math

import numpy as np
import torch

import tensorflow as tf

_QUEUE: list = []


def push_uniforms(*arrays):
    """Queue uniform(0,1) arrays; every sampler call below consumes exactly one, first in first out."""
    for a in arrays:
        _QUEUE.append(a if isinstance(a, torch.Tensor) else torch.as_tensor(np.asarray(a)))


def pending_uniforms():
    return len(_QUEUE)


def _pop(shape, dtype):
    if not _QUEUE:
        raise RuntimeError("sampler called with no queued uniforms (push_uniforms first)")
    u = _QUEUE.pop(0)
    if list(u.shape) != list(shape):
        raise RuntimeError("queued uniforms have shape %s, the reference draws %s" % (list(u.shape), list(shape)))
    return u.to(dtype)


def _t(v, dtype=None):
    if isinstance(v, tf.Tensor):
        return v.t
    if isinstance(v, torch.Tensor):
        return v
    return torch.tensor(v, dtype=dtype or torch.float32)


class Distribution:
    def __init__(self, dtype=None, reparameterization_type=None, validate_args=False, allow_nan_stats=True,
                 parameters=None, name=None):
        self._dtype = dtype
        self._reparameterization_type = reparameterization_type

    @property
    def dtype(self):
        return self._dtype

    def log_prob(self, value, **kwargs):
        return self._log_prob(tf.convert_to_tensor(value), **kwargs)

    def prob(self, value, **kwargs):
        return tf.exp(self.log_prob(value, **kwargs))

    def mean(self, **kwargs):
        return self._mean(**kwargs)

    def sample(self, sample_shape=(), seed=None, **kwargs):
        if isinstance(sample_shape, (int, np.integer)):
            shape = [int(sample_shape)]
        else:
            shape = [int(s) for s in sample_shape]
        n = int(math.prod(shape)) if shape else 1
        out = self._sample_n(n, seed=seed, **kwargs)
        rest = list(out.shape)[1:]
        return tf.reshape(out, shape + rest)


class Normal(Distribution):
    def __init__(self, loc, scale, **kw):
        self.loc, self.scale = tf.convert_to_tensor(loc), tf.convert_to_tensor(scale)
        super().__init__(dtype=self.loc.dtype)

    def _log_prob(self, x):
        # normal.py: -0.5 * squared_difference(x / scale, loc / scale) - (0.5 * log(2 pi) + log(scale))
        xs, ls = x.t / self.scale.t, self.loc.t / self.scale.t
        return tf.Tensor(-0.5 * (xs - ls) ** 2 - (0.5 * math.log(2.0 * math.pi) + torch.log(self.scale.t)))

    def _sample_n(self, n, seed=None):
        raise RuntimeError("Normal.sample is outside the tested path (latents are inputs of the fixtures)")


class Logistic(Distribution):
    def __init__(self, loc, scale, **kw):
        self.loc, self.scale = tf.convert_to_tensor(loc), tf.convert_to_tensor(scale)
        super().__init__(dtype=self.loc.dtype)

    def _sample_n(self, n, seed=None):
        bshape = list(torch.broadcast_shapes(self.loc.t.shape, self.scale.t.shape))
        u = _pop([n] + bshape, self.loc.t.dtype)
        sampled = torch.log(u) - torch.log1p(-u)
        return tf.Tensor(sampled * self.scale.t + self.loc.t)


def _gumbel_argmax(logits, u):
    return torch.argmax(logits - torch.log(-torch.log(u)), dim=-1)


class Categorical(Distribution):
    def __init__(self, logits=None, dtype=tf.int32, **kw):
        self.logits = tf.convert_to_tensor(logits)
        super().__init__(dtype=dtype)

    def _sample_n(self, n, seed=None):
        lg = self.logits.t
        u = _pop([n] + list(lg.shape), lg.dtype)
        return tf.Tensor(_gumbel_argmax(lg, u))


class OneHotCategorical(Distribution):
    def __init__(self, logits=None, dtype=tf.int32, **kw):
        self.logits = tf.convert_to_tensor(logits)
        super().__init__(dtype=dtype)

    def _sample_n(self, n, seed=None):
        lg = self.logits.t
        u = _pop([n] + list(lg.shape), lg.dtype)
        idx = _gumbel_argmax(lg, u)
        return tf.Tensor(torch.nn.functional.one_hot(idx, lg.shape[-1]).to(self.dtype))
