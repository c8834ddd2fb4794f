"""Minimal stand-in for the `tensorflow` symbols the reference's observation-model path uses, on torch-CPU tensors.

TEST INFRASTRUCTURE ONLY -- see oracle/tf_shim/README.md.  Not a TensorFlow re-implementation: exactly the ops that
utils/mdl*.py, utils/discretized_logistic.py, utils/utils.py and models/loss.py of nbip/vae-mdl call, each with
TensorFlow's semantics (dtype-preserving python scalars, `tf.where` gradient by masking, `tf.maximum` tie rule,
softplus thresholds).
"""
from __future__ import annotations

import math as _math
import types as _types

import numpy as _np
import torch as _torch

__version__ = "2.8-shim"

float32 = _torch.float32
float64 = _torch.float64
int32 = _torch.int32
int64 = _torch.int64


class TensorShape(list):
    """`tf.TensorShape` for static shapes: a list that survives `shape[:-1] + [3, 3 * n]` (utils/mdl.py:102)."""

    def __add__(self, other):
        return TensorShape(list(self) + list(other))

    def __radd__(self, other):
        return TensorShape(list(other) + list(self))

    def __getitem__(self, i):
        r = list.__getitem__(self, i)
        return TensorShape(r) if isinstance(i, slice) else r

    def as_list(self):
        return list(self)


def _raw(v):
    return v.t if isinstance(v, Tensor) else v


def _wrap(t):
    return Tensor(t) if isinstance(t, _torch.Tensor) else t


class Tensor:
    """Eager tensor: a torch tensor behind the few attributes / operators the reference touches."""

    __array_ufunc__ = None  # numpy scalars defer to our reflected operators

    def __init__(self, t):
        self.t = t

    # -- static information -------------------------------------------------
    @property
    def shape(self):
        return TensorShape(self.t.shape)

    def get_shape(self):
        return self.shape

    @property
    def dtype(self):
        return self.t.dtype

    def numpy(self):
        return self.t.detach().numpy()

    def __len__(self):
        return self.t.shape[0]

    def __repr__(self):
        return "tf_shim.Tensor(%r)" % (self.t,)

    def __getitem__(self, idx):
        if isinstance(idx, tuple):
            idx = tuple(_raw(i) for i in idx)
        else:
            idx = _raw(idx)
        return Tensor(self.t[idx])

    # -- arithmetic (python / numpy scalars never promote the dtype, like TF's weak constants) ------------------
    @staticmethod
    def _other(self_t, o):
        o = _raw(o)
        if isinstance(o, _torch.Tensor):
            if o.dtype != self_t.dtype and o.is_floating_point() and self_t.is_floating_point():
                # TF raises on a float32/float64 mix; the one place the reference mixes them is the float32 cast of
                # log(interval_width) (utils/mdl.py:163) -- keep that rounding and carry on in the wider type.
                return o.to(_torch.promote_types(o.dtype, self_t.dtype))
            return o
        if isinstance(o, (_np.generic,)):
            return o.item()
        return o

    def __add__(self, o): return Tensor(self.t + self._other(self.t, o))
    def __radd__(self, o): return Tensor(self._other(self.t, o) + self.t)
    def __sub__(self, o): return Tensor(self.t - self._other(self.t, o))
    def __rsub__(self, o): return Tensor(self._other(self.t, o) - self.t)
    def __mul__(self, o): return Tensor(self.t * self._other(self.t, o))
    def __rmul__(self, o): return Tensor(self._other(self.t, o) * self.t)
    def __truediv__(self, o): return Tensor(self.t / self._other(self.t, o))
    def __rtruediv__(self, o): return Tensor(self._other(self.t, o) / self.t)
    def __neg__(self): return Tensor(-self.t)
    def __lt__(self, o): return Tensor(self.t < self._other(self.t, o))
    def __le__(self, o): return Tensor(self.t <= self._other(self.t, o))
    def __gt__(self, o): return Tensor(self.t > self._other(self.t, o))
    def __ge__(self, o): return Tensor(self.t >= self._other(self.t, o))


def _axes(axis):
    if axis is None:
        return None
    if isinstance(axis, (list, tuple)):
        return tuple(int(a) for a in axis)
    return int(axis)


# ---------------------------------------------------------------------------------------------------------------
# construction / shape ops
# ---------------------------------------------------------------------------------------------------------------
def convert_to_tensor(v, dtype=None):
    if isinstance(v, Tensor):
        return v if dtype is None else Tensor(v.t.to(dtype))
    if isinstance(v, _torch.Tensor):
        return Tensor(v if dtype is None else v.to(dtype))
    a = _np.asarray(v)
    if dtype is None and a.dtype == _np.float64 and not isinstance(v, _np.ndarray):
        dtype = float32  # python floats become float32 constants in TF
    t = _torch.from_numpy(_np.ascontiguousarray(a))
    return Tensor(t if dtype is None else t.to(dtype))


constant = convert_to_tensor


def cast(v, dtype):
    if isinstance(v, Tensor):
        return Tensor(v.t.to(dtype))
    return Tensor(_torch.tensor(v, dtype=dtype))


def zeros(shape, dtype=float32):
    return Tensor(_torch.zeros(list(shape), dtype=dtype))


def reshape(x, shape):
    return Tensor(_raw(x).reshape([int(s) for s in shape]))


def expand_dims(x, axis):
    return Tensor(_raw(x).unsqueeze(int(axis)))


def concat(values, axis):
    return Tensor(_torch.cat([_raw(v) for v in values], dim=int(axis)))


def split(value, num_or_size_splits, axis=0):
    t = _raw(value)
    assert isinstance(num_or_size_splits, int)
    assert t.shape[axis] % num_or_size_splits == 0
    return [Tensor(p) for p in _torch.split(t, t.shape[axis] // num_or_size_splits, dim=axis)]


def repeat(x, repeats, axis):
    return Tensor(_torch.repeat_interleave(_raw(x), int(repeats), dim=int(axis)))


def one_hot(indices, depth, axis=-1, dtype=float32):
    assert axis == -1
    return Tensor(_torch.nn.functional.one_hot(_raw(indices).long(), int(depth)).to(dtype))


def argmax(x, axis=None):
    return Tensor(_torch.argmax(_raw(x), dim=int(axis)))


# ---------------------------------------------------------------------------------------------------------------
# elementwise
# ---------------------------------------------------------------------------------------------------------------
def exp(x):
    return Tensor(_torch.exp(_raw(x)))


def _log(x):
    if isinstance(x, Tensor):
        return Tensor(_torch.log(x.t))
    # tf.math.log(python float) -> float32 scalar tensor (utils/mdl.py:163 then casts it to float32 again)
    return Tensor(_torch.log(_torch.tensor(float(x), dtype=_torch.float32)))


def where(cond, a, b):
    c = _raw(cond)
    a, b = _raw(a), _raw(b)
    return Tensor(_torch.where(c, a, b))


def _max_with_tf_ties(x, y):
    """tf.maximum: value max(x, y); gradient to x where x >= y, to y elsewhere (math_grad._MaximumMinimumGrad)."""
    x, y = _raw(x), _raw(y)
    assert isinstance(x, _torch.Tensor), "the reference always passes the tensor first"
    if not isinstance(y, _torch.Tensor):
        y = _torch.full_like(x, float(y))
    return _torch.where(x >= y, x, y)


def maximum(x, y):
    return Tensor(_max_with_tf_ties(x, y))


def minimum(x, y):
    x, y = _raw(x), _raw(y)
    if not isinstance(y, _torch.Tensor):
        y = _torch.full_like(x, float(y))
    return Tensor(_torch.where(x <= y, x, y))  # gradient to x where x <= y


def clip_by_value(x, lo, hi):
    return minimum(maximum(x, lo), hi)


def less_equal(a, b):
    return Tensor(_raw(a) <= _raw(b))


def greater_equal(a, b):
    return Tensor(_raw(a) >= _raw(b))


def greater(a, b):
    return Tensor(_raw(a) > _raw(b))


class _Softplus(_torch.autograd.Function):
    """tf.nn.softplus (tensorflow/core/kernels/softplus_op.h): threshold = log(eps) + 2;
    x > -threshold -> x; x < threshold -> exp(x); else log1p(exp(x)).  SoftplusGrad: g / (exp(-x) + 1)."""

    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        thr = _math.log(_torch.finfo(x.dtype).eps) + 2.0
        too_large = x > -thr
        too_small = x < thr
        e = _torch.exp(_torch.where(too_large, _torch.zeros_like(x), x))
        return _torch.where(too_large, x, _torch.where(too_small, e, _torch.log1p(e)))

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        return g / (_torch.exp(-x) + 1.0)


def _softplus(x):
    return Tensor(_Softplus.apply(_raw(x)))


def _sigmoid(x):
    return Tensor(_torch.sigmoid(_raw(x)))


def _tanh(x):
    return Tensor(_torch.tanh(_raw(x)))


# ---------------------------------------------------------------------------------------------------------------
# reductions
# ---------------------------------------------------------------------------------------------------------------
def reduce_sum(x, axis=None, keepdims=False):
    t = _raw(x)
    ax = _axes(axis)
    return Tensor(t.sum() if ax is None else t.sum(dim=ax, keepdim=keepdims))


def reduce_mean(x, axis=None, keepdims=False):
    t = _raw(x)
    ax = _axes(axis)
    return Tensor(t.mean() if ax is None else t.mean(dim=ax, keepdim=keepdims))


def reduce_max(x, axis=None, keepdims=False):
    t = _raw(x)
    ax = _axes(axis)
    return Tensor(t.max() if ax is None else _torch.amax(t, dim=ax, keepdim=keepdims))  # ties share the gradient


def reduce_logsumexp(x, axis=None, keepdims=False):
    """math_ops.reduce_logsumexp: log(sum(exp(x - stop_gradient(max)))) + max, max := 0 where it is not finite."""
    t = _raw(x)
    ax = _axes(axis)
    m = _torch.amax(t, dim=ax, keepdim=True).detach()
    m = _torch.where(_torch.isfinite(m), m, _torch.zeros_like(m))
    r = _torch.log(_torch.sum(_torch.exp(t - m), dim=ax, keepdim=True)) + m
    return Tensor(r if keepdims else r.squeeze(ax))


def _log_softmax(x, axis=-1):
    return Tensor(_torch.log_softmax(_raw(x), dim=int(axis)))


def _reduce_prod(v, axis=None):
    if isinstance(v, Tensor):
        return Tensor(v.t.prod())
    return Tensor(_torch.tensor(int(_np.prod(list(v))), dtype=_torch.int32))


def _uniform(shape, minval=0.0, maxval=1.0, dtype=float32, seed=None):
    raise RuntimeError("tf.random.uniform is not on the tested path (only bernoullisample, utils/utils.py:14-17, calls it)")


nn = _types.SimpleNamespace(softplus=_softplus, sigmoid=_sigmoid, tanh=_tanh, log_softmax=_log_softmax)
math = _types.SimpleNamespace(log=_log, maximum=maximum, minimum=minimum, log_softmax=_log_softmax, greater=greater,
                              reduce_prod=_reduce_prod, exp=exp, softplus=_softplus, sigmoid=_sigmoid, tanh=_tanh)
random = _types.SimpleNamespace(uniform=_uniform)
log = _log


from . import keras  # noqa: E402,F401  (class names only; see keras/__init__.py)


def function(fn=None, **_kw):  # @tf.function: eager here
    return fn if fn is not None else (lambda f: f)
