"""Op-for-op CPU restatement (torch-CPU tensors) of the reference's observation-model path.

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.  Pinned against the
reference's own source executed over ``oracle/tf_shim`` (tests/golden/refsrc_*.npz,
tests/test_reference_golden.py); TensorFlow itself cannot run here, so TF's
float32 kernel rounding stays unpinned.

Two flavours of every function, selected by the dtype of the inputs:

* float64 ("ref64")  -- ground truth for the parity tests.
* float32 ("ref32")  -- same op order as the TF graph, TF's softplus thresholds
  and ``reduce_logsumexp`` formulation; stands in for "the TF CPU path" and is
  what ``bench.py`` times as the CPU baseline (``cpu_baseline.kind = "port"``).

All ``file:line`` citations are relative to the reference repository root
(nbip/vae-mdl).  Gradients come from torch autograd over these restatements,
exactly as the reference gets them from ``tf.GradientTape``
(models/model05.py:141-145).
"""
from __future__ import annotations

import math

import torch

__all__ = [
    "softplus",
    "log_sum_exp",
    "log_prob_from_logits",
    "discretized_mix_logistic_loss",
    "sample_from_discretized_mix_logistic",
    "modl_split_params",
    "modl_log_prob",
    "modl_openai_log_prob",
    "modl_openai_iwae_log_prob",
    "modl_sample_mdl",
    "mdl_plain_get_mixture_params",
    "mdl_plain_log_prob",
    "mdl_plain_sample",
    "modl_openai_sample",
    "modl_openai_iwae_sample",
    "dlogistic_log_prob",
    "dlogistic_sample",
    "logmeanexp",
    "iwae_loss",
    "elbo_loss",
    "model06_loss",
    "quantise",
    "normalize_u8",
    "gumbel_argmax",
    "logistic_eps",
]


# --------------------------------------------------------------------------- #
# TensorFlow primitive semantics                                               #
# --------------------------------------------------------------------------- #
def softplus(x: torch.Tensor) -> torch.Tensor:
    """``tf.nn.softplus``.

    float32: TF's kernel (tensorflow/core/kernels/softplus_op.h) uses
    ``threshold = log(eps) + 2``: ``x > -threshold -> x``; ``x < threshold ->
    exp(x)``; else ``log1p(exp(x))``  [TF-internal, stated from knowledge of TF].
    float64: the exact, overflow-free form ``max(x,0) + log1p(exp(-|x|))``.
    """
    if x.dtype == torch.float32:
        thr = math.log(torch.finfo(torch.float32).eps) + 2.0
        too_large = x > -thr
        too_small = x < thr
        xe = torch.exp(torch.where(too_large, torch.zeros_like(x), x))
        return torch.where(too_large, x, torch.where(too_small, xe, torch.log1p(xe)))
    return torch.clamp(x, min=0) + torch.log1p(torch.exp(-torch.abs(x)))


def _reduce_logsumexp(x: torch.Tensor, dim: int) -> torch.Tensor:
    """``tf.reduce_logsumexp``: ``log(sum(exp(x - stop_gradient(max)))) + max``  [TF-internal]."""
    m = torch.amax(x, dim=dim, keepdim=True).detach()
    m = torch.where(torch.isfinite(m), m, torch.zeros_like(m))
    return torch.log(torch.sum(torch.exp(x - m), dim=dim)) + m.squeeze(dim)


def _log_softmax(x: torch.Tensor, dim: int) -> torch.Tensor:
    """``tf.nn.log_softmax``: ``x - logsumexp(x)``  [TF-internal]."""
    return x - _reduce_logsumexp(x, dim).unsqueeze(dim)


def _tf_maximum_const(x: torch.Tensor, c: float) -> torch.Tensor:
    """``tf.maximum(x, c)``; gradient goes to ``x`` when ``x >= c``  [TF-internal]."""
    return torch.where(x >= c, x, torch.full_like(x, c))


# --------------------------------------------------------------------------- #
# utils/mdl_openai.py  (byte-identical copy in utils/mdl_openai_iwae.py)       #
# --------------------------------------------------------------------------- #
def log_sum_exp(x: torch.Tensor) -> torch.Tensor:
    """utils/mdl_openai.py:68-73."""
    m = torch.amax(x, dim=-1)
    m2 = torch.amax(x, dim=-1, keepdim=True)
    return m + torch.log(torch.sum(torch.exp(x - m2), dim=-1))


def log_prob_from_logits(x: torch.Tensor) -> torch.Tensor:
    """utils/mdl_openai.py:76-80."""
    m = torch.amax(x, dim=-1, keepdim=True)
    return x - m - torch.log(torch.sum(torch.exp(x - m), dim=-1, keepdim=True))


def discretized_mix_logistic_loss(x: torch.Tensor, l: torch.Tensor, sum_all: bool = True) -> torch.Tensor:
    """utils/mdl_openai.py:83-157.  ``x`` in [-1,1] ``[N,H,W,3]``, ``l`` ``[N,H,W,10*M]``."""
    xs = list(x.shape)
    ls = list(l.shape)
    nr_mix = ls[-1] // 10
    logit_probs = l[:, :, :, :nr_mix]                                           # :90
    l = l[:, :, :, nr_mix:].reshape(xs + [nr_mix * 3])                          # :91
    means = l[:, :, :, :, :nr_mix]                                              # :92
    log_scales = _tf_maximum_const(l[:, :, :, :, nr_mix:2 * nr_mix], -7.0)      # :93
    coeffs = torch.tanh(l[:, :, :, :, 2 * nr_mix:3 * nr_mix])                   # :94
    x = x.reshape(xs + [1]) + torch.zeros(xs + [nr_mix], dtype=x.dtype)         # :95-97
    m2 = (means[:, :, :, 1, :] + coeffs[:, :, :, 0, :] * x[:, :, :, 0, :]).reshape(xs[0], xs[1], xs[2], 1, nr_mix)
    m3 = (
        means[:, :, :, 2, :]
        + coeffs[:, :, :, 1, :] * x[:, :, :, 0, :]
        + coeffs[:, :, :, 2, :] * x[:, :, :, 1, :]
    ).reshape(xs[0], xs[1], xs[2], 1, nr_mix)                                   # :98-107
    means = torch.cat([means[:, :, :, 0, :].reshape(xs[0], xs[1], xs[2], 1, nr_mix), m2, m3], dim=3)
    centered_x = x - means                                                      # :111
    inv_stdv = torch.exp(-log_scales)                                           # :112
    plus_in = inv_stdv * (centered_x + 1.0 / 255.0)                             # :113
    cdf_plus = torch.sigmoid(plus_in)
    min_in = inv_stdv * (centered_x - 1.0 / 255.0)                              # :115
    cdf_min = torch.sigmoid(min_in)
    log_cdf_plus = plus_in - softplus(plus_in)                                  # :117
    log_one_minus_cdf_min = -softplus(min_in)                                   # :120
    cdf_delta = cdf_plus - cdf_min                                              # :123
    mid_in = inv_stdv * centered_x                                              # :124
    log_pdf_mid = mid_in - log_scales - 2.0 * softplus(mid_in)                  # :125-127
    log_probs = torch.where(
        x < -0.999,
        log_cdf_plus,
        torch.where(
            x > 0.999,
            log_one_minus_cdf_min,
            torch.where(
                cdf_delta > 1e-5,
                torch.log(torch.clamp(cdf_delta, min=1e-12)),
                log_pdf_mid - math.log(127.5),
            ),
        ),
    )                                                                           # :138-150
    log_probs = torch.sum(log_probs, dim=3) + log_prob_from_logits(logit_probs)  # :152
    if sum_all:
        return -torch.sum(log_sum_exp(log_probs))                               # :154
    return log_sum_exp(log_probs)                                               # :157


def gumbel_argmax(logit_probs: torch.Tensor, u_mix: torch.Tensor) -> torch.Tensor:
    """Explicit-noise mixture selection, utils/mdl_openai.py:167 (commented PixelCNN++ original):
    ``argmax_m(logit_m - log(-log u_m))``.  Computed in float64; first maximum wins ties."""
    g = logit_probs.double() - torch.log(-torch.log(u_mix.double()))
    return torch.argmax(g, dim=-1)


def logistic_eps(u: torch.Tensor) -> torch.Tensor:
    """Explicit-noise standard-logistic draw, utils/mdl_openai.py:185-186: ``log u - log(1-u)`` (float64)."""
    u = u.double()
    return torch.log(u) - torch.log(1.0 - u)


def sample_from_discretized_mix_logistic(l: torch.Tensor, nr_mix: int, u_mix: torch.Tensor, u_log: torch.Tensor):
    """utils/mdl_openai.py:160-193 with the explicit-noise lines :167 and :185-186.

    ``l [N,H,W,10M]``, ``u_mix [N,H,W,M]``, ``u_log [N,H,W,3]`` (uniforms in (0,1)).
    All arithmetic in float64 (the reference has no explicit-noise code path to
    match bit-for-bit; float64 on both sides makes index / quantised-value
    equality a well-posed test).  Returns ``(x [N,H,W,3] float64 in [-1,1], idx [N,H,W] int64)``.
    """
    ls = list(l.shape)
    xs = ls[:-1] + [3]
    l = l.double()
    logit_probs = l[:, :, :, :nr_mix]                                            # :164
    l = l[:, :, :, nr_mix:].reshape(xs + [nr_mix * 3])                           # :165
    idx = gumbel_argmax(logit_probs, u_mix)                                      # :167
    sel = torch.nn.functional.one_hot(idx, nr_mix).to(l.dtype).unsqueeze(-2)     # :168
    means = torch.sum(l[:, :, :, :, :nr_mix] * sel, dim=4)                       # :177
    log_scales = torch.clamp(torch.sum(l[:, :, :, :, nr_mix:2 * nr_mix] * sel, dim=4), min=-7.0)  # :178-180
    coeffs = torch.sum(torch.tanh(l[:, :, :, :, 2 * nr_mix:3 * nr_mix]) * sel, dim=4)            # :181
    x = means + torch.exp(log_scales) * logistic_eps(u_log)                      # :185-186
    x0 = torch.clamp(x[:, :, :, 0], -1.0, 1.0)                                   # :190
    x1 = torch.clamp(x[:, :, :, 1] + coeffs[:, :, :, 0] * x0, -1.0, 1.0)         # :191
    x2 = torch.clamp(x[:, :, :, 2] + coeffs[:, :, :, 1] * x0 + coeffs[:, :, :, 2] * x1, -1.0, 1.0)  # :192
    return torch.stack([x0, x1, x2], dim=3), idx                                 # :193


# --------------------------------------------------------------------------- #
# utils/mdl.py : MixtureDiscretizedLogistic                                     #
# --------------------------------------------------------------------------- #
def modl_split_params(parameters: torch.Tensor):
    """utils/mdl.py:94-112 (``_split_params``)."""
    n_mix = parameters.shape[-1] // 10
    mix_logits = parameters[..., :n_mix]                                         # :98
    rest = parameters[..., n_mix:].reshape(list(parameters.shape[:-1]) + [3, 3 * n_mix])  # :101-103
    _loc, logscale, coeffs = torch.split(rest, n_mix, dim=-1)                    # :106-108
    logscale = _tf_maximum_const(logscale, -7.0)                                 # :109
    coeffs = torch.tanh(coeffs)                                                  # :110
    return _loc, logscale, coeffs, mix_logits


def _mdl_autoregressive_params(parameters: torch.Tensor, x: torch.Tensor):
    """utils/mdl.py:114-151.  ``x`` already in [-1,1]."""
    _loc, logscale, coeffs, mix_logits = modl_split_params(parameters)
    loc_r = _loc[..., 0, :]                                                      # :139
    loc_g = _loc[..., 1, :] + coeffs[..., 0, :] * x[..., 0, None]                # :140
    loc_b = _loc[..., 2, :] + coeffs[..., 1, :] * x[..., 0, None] + coeffs[..., 2, :] * x[..., 1, None]  # :141-145
    loc = torch.cat([loc_r[..., None, :], loc_g[..., None, :], loc_b[..., None, :]], dim=-2)  # :147-149
    return loc, logscale, mix_logits


def _discretized_logistic_log_prob(x, loc, logscale, dx, low, high, interval_width, approx_divides=False):
    """utils/mdl.py:157-207 and utils/discretized_logistic.py:27-78 (same body).

    ``approx_divides``: the plain class divides by ``exp(logscale)`` in the
    approximate branch (utils/discretized_logistic.py:31) where the mixture
    class multiplies by ``exp(-logscale)`` (utils/mdl.py:161).
    """
    centered_x = x - loc                                                         # mdl.py:166
    inv_std = torch.exp(-logscale)                                               # :167
    interval_start = (centered_x - dx) * inv_std                                 # :172
    interval_stop = (centered_x + dx) * inv_std                                  # :173
    prob = torch.sigmoid(interval_stop) - torch.sigmoid(interval_start)          # :176
    prob = torch.clamp(prob, min=1e-12)                                          # :180
    left_edge = interval_stop - softplus(interval_stop)                          # :185
    right_edge = -softplus(interval_start)                                       # :186
    if approx_divides:
        a = (x - loc) / torch.exp(logscale)                                      # discretized_logistic.py:31
    else:
        a = (x - loc) * torch.exp(-logscale)                                     # mdl.py:161
    log_pdf_val = -a - logscale - 2 * softplus(-a)                               # :162
    log_prob_approx = log_pdf_val + math.log(interval_width)                     # :163
    safe_log_prob = torch.where(prob > 1e-5, torch.log(prob), log_prob_approx)   # :193
    with_left = torch.where(x <= low, left_edge, safe_log_prob)                  # :200-202
    return torch.where(x >= high, right_edge, with_left)                         # :203-205


def modl_log_prob(parameters: torch.Tensor, x01: torch.Tensor) -> torch.Tensor:
    """``MixtureDiscretizedLogistic.log_prob`` -- utils/mdl.py:56-92.

    ``parameters [..., B, H, W, 10M]``; ``x01`` in [0,1], ``[B,H,W,3]`` or any
    shape broadcastable against the leading dims (models/model05.py:173 passes
    ``[H,W,3]``).  Returns ``[..., B, H, W, 1]``.
    """
    x = x01 * 2.0 - 1.0                                                          # :65
    loc, logscale, mix_logits = _mdl_autoregressive_params(parameters, x)        # :68
    lp = _discretized_logistic_log_prob(
        x[..., None], loc, logscale, dx=(2.0 / 255.0) / 2.0, low=-1.0, high=1.0, interval_width=2.0 / 255.0
    )                                                                            # :72-74, :47-52
    mix_log_weights = _log_softmax(mix_logits, -1)                               # :78
    weighted = torch.sum(lp, dim=-2) + mix_log_weights                           # :83-85
    out = _reduce_logsumexp(weighted, -1)                                        # :89
    return out.unsqueeze(-1)                                                     # :92


def modl_sample_mdl(parameters: torch.Tensor, u_mix: torch.Tensor, u_log: torch.Tensor):
    """``MixtureDiscretizedLogistic._sample_n`` -- utils/mdl.py:209-252 with explicit noise.

    ``u_mix [..., H, W, M]`` (Gumbel-argmax selection, see ``gumbel_argmax``),
    ``u_log [..., H, W, 3, M]`` (one logistic draw for EVERY mixture, :213).
    float64 internally.  Returns ``(x01 [..., H, W, 3] float64 in [0,1], idx int64)``.
    """
    _loc, logscale, coeffs, mix_logits = modl_split_params(parameters.double())  # :210
    logistic = _loc + torch.exp(logscale) * logistic_eps(u_log)                  # :213
    sample_r = torch.clamp(logistic[..., 0, :], -1.0, 1.0)                       # :218
    sample_g = torch.clamp(logistic[..., 1, :] + coeffs[..., 0, :] * sample_r, -1.0, 1.0)  # :219-221
    sample_b = torch.clamp(
        logistic[..., 2, :] + coeffs[..., 1, :] * sample_r + coeffs[..., 2, :] * sample_g, -1.0, 1.0
    )                                                                            # :222-228
    ar = torch.cat([sample_r[..., None, :], sample_g[..., None, :], sample_b[..., None, :]], dim=-2)
    idx = gumbel_argmax(mix_logits, u_mix)                                       # :236-238
    onehot = torch.nn.functional.one_hot(idx, mix_logits.shape[-1]).to(ar.dtype).unsqueeze(-2)  # :240
    selected = torch.sum(ar * onehot, dim=-1)                                    # :245-247
    return selected * 0.5 + 0.5, idx                                             # :250


# --------------------------------------------------------------------------- #
# utils/mdl_plain.py : PixelMixtureDiscretizedLogistic (no conditioning on x)    #
# --------------------------------------------------------------------------- #
def mdl_plain_get_mixture_params(parameters: torch.Tensor):
    """``get_mixture_params`` -- utils/mdl_plain.py:124-168: the green / blue means are chained on the component's OWN
    red / green means (:160-162), not on the observed x.  Returns ``loc, logscale [..., 3, M]``, ``mix_logits [..., M]``."""
    _loc, logscale, coeffs, mix_logits = modl_split_params(parameters)           # :143-154 (same split, clamp, tanh)
    loc_r = _loc[..., 0, :]                                                      # :160
    loc_g = _loc[..., 1, :] + coeffs[..., 0, :] * loc_r                          # :161
    loc_b = _loc[..., 2, :] + coeffs[..., 1, :] * loc_r + coeffs[..., 2, :] * loc_g  # :162
    loc = torch.cat([loc_r[..., None, :], loc_g[..., None, :], loc_b[..., None, :]], dim=-2)  # :164-166
    return loc, logscale, mix_logits


def mdl_plain_log_prob(parameters: torch.Tensor, x01: torch.Tensor, low=-1.0, high=1.0, levels=256.0) -> torch.Tensor:
    """``PixelMixtureDiscretizedLogistic.log_prob`` -- utils/mdl_plain.py:36-66; ``low, high, levels`` are the constructor
    arguments (:18) handed to the ``DiscretizedLogistic`` base (:28-30).
    Returns ``[..., H, W]`` (no trailing 1: :66 reduces over the mixture axis only)."""
    loc, logscale, mix_logits = mdl_plain_get_mixture_params(parameters)         # :27
    x = x01 * 2.0 - 1.0                                                          # :45
    lp = dlogistic_log_prob(x[..., None], loc, logscale, low, high, levels)      # :49-51 (DiscretizedLogistic.log_prob)
    mix_log_weights = _log_softmax(mix_logits, -1)                               # :55
    weighted = torch.sum(lp, dim=-2) + mix_log_weights                           # :59-61
    return _reduce_logsumexp(weighted, -1)                                       # :66


def mdl_plain_sample(parameters: torch.Tensor, u_mix: torch.Tensor, u_log, low=-1.0, high=1.0):
    """``PixelMixtureDiscretizedLogistic.sample`` (utils/mdl_plain.py:68-102) / ``.mean`` (:104-121) with explicit noise:
    ``u_mix [..., H, W, M]`` selects the component (Gumbel-argmax stands in for tfd.Categorical), ``u_log
    [..., H, W, 3, M]`` drives ``DiscretizedLogistic.sample`` for every component (:86-88); ``u_log=None`` gives
    ``mean()``: the selected, clipped locations.  float64.  Returns ``(x01 [..., H, W, 3], idx)``."""
    loc, logscale, mix_logits = mdl_plain_get_mixture_params(parameters.double())
    idx = gumbel_argmax(mix_logits, u_mix)                                       # :77-78 / :107-108
    onehot = torch.nn.functional.one_hot(idx, mix_logits.shape[-1]).to(loc.dtype).unsqueeze(-2)  # :79-81
    if u_log is None:
        vals = loc                                                               # :116-118
    else:
        vals = dlogistic_sample(loc, logscale, u_log, low, high)                 # :86-88 (clipped to [low, high])
    sel = torch.sum(vals * onehot, dim=-1)                                       # :93-95
    if u_log is None:
        sel = torch.clamp(sel, -1.0, 1.0)                                        # :119
    return (sel + 1.0) / 2.0, idx                                                # :98 / :120


# --------------------------------------------------------------------------- #
# wrappers: utils/mdl_openai.py:15-58, utils/mdl_openai_iwae.py:16-102           #
# --------------------------------------------------------------------------- #
def modl_openai_log_prob(logits: torch.Tensor, x_pm1: torch.Tensor) -> torch.Tensor:
    """``MixtureDiscretizedLogisticOpenai._log_prob`` -- utils/mdl_openai.py:31-32.  -> ``[B,H,W]``."""
    return discretized_mix_logistic_loss(x_pm1, logits, sum_all=False)


def modl_openai_sample(logits: torch.Tensor, n: int, u_mix: torch.Tensor, u_log: torch.Tensor):
    """``MixtureDiscretizedLogisticOpenai._sample_n`` -- utils/mdl_openai.py:34-55.

    ``u_mix [n*B,H,W,M]``, ``u_log [n*B,H,W,3]`` (n-major like the tiled logits, :39-45).
    Returns ``(x [n,B,H,W,3] in [-1,1], idx [n,B,H,W])``.
    """
    n_logits = logits.unsqueeze(0).repeat_interleave(n, dim=0)                   # :39
    n_logits = n_logits.reshape([logits.shape[0] * n] + list(logits.shape[1:]))  # :43-45
    x, idx = sample_from_discretized_mix_logistic(n_logits, logits.shape[-1] // 10, u_mix, u_log)  # :49
    return x.reshape([n] + list(logits.shape[:-1]) + [3]), idx.reshape([n] + list(logits.shape[:-1]))  # :53


def modl_openai_iwae_log_prob(logits: torch.Tensor, x01: torch.Tensor) -> torch.Tensor:
    """``MixtureDiscretizedLogisticOpenaiIWAE._log_prob`` -- utils/mdl_openai_iwae.py:33-67."""
    x = x01 * 2.0 - 1.0                                                          # :35
    shape = list(logits.shape)
    logits_reshaped = logits.reshape([-1] + shape[-3:])                          # :38
    repeats = logits_reshaped.shape[0] // x.shape[0]                             # :49
    x_repeat = x[None].repeat_interleave(repeats, dim=0)                         # :55
    x_reshaped = x_repeat.reshape([-1] + list(x.shape[-3:]))                     # :56
    lp = discretized_mix_logistic_loss(x_reshaped, logits_reshaped, sum_all=False)  # :60
    return lp.reshape(shape[:-1]).unsqueeze(-1)                                  # :64-67


def modl_openai_iwae_sample(logits: torch.Tensor, n: int, u_mix: torch.Tensor, u_log: torch.Tensor):
    """``MixtureDiscretizedLogisticOpenaiIWAE._sample_n`` -- utils/mdl_openai_iwae.py:69-99.  -> [0,1]."""
    shape = list(logits.shape)
    logits_reshaped = logits.reshape([-1] + shape[-3:])                          # :74
    n_logits = logits_reshaped.unsqueeze(0).repeat_interleave(n, dim=0)          # :78
    n_logits = n_logits.reshape([logits_reshaped.shape[0] * n] + shape[-3:])     # :82-84
    x, idx = sample_from_discretized_mix_logistic(n_logits, shape[-1] // 10, u_mix, u_log)  # :88
    x = x.reshape([n] + shape[:-1] + [3])                                        # :92-97
    return x * 0.5 + 0.5, idx.reshape([n] + shape[:-1])                          # :99


# --------------------------------------------------------------------------- #
# utils/discretized_logistic.py : DiscretizedLogistic                           #
# --------------------------------------------------------------------------- #
def dlogistic_log_prob(x, loc, logscale, low=-1.0, high=1.0, levels=256.0) -> torch.Tensor:
    """``DiscretizedLogistic.log_prob`` -- utils/discretized_logistic.py:10-21, :35-78."""
    interval_width = (high - low) / (levels - 1.0)                               # :18
    dx = interval_width / 2.0                                                    # :21
    return _discretized_logistic_log_prob(x, loc, logscale, dx, low, high, interval_width, approx_divides=True)


def dlogistic_sample(loc, logscale, u, low=-1.0, high=1.0) -> torch.Tensor:
    """``DiscretizedLogistic.sample`` -- utils/discretized_logistic.py:80-85, explicit noise, float64."""
    x = loc.double() + torch.exp(logscale.double()) * logistic_eps(u)            # :81-82
    return torch.clamp(x, low, high)                                             # :83


# --------------------------------------------------------------------------- #
# utils/utils.py:9-11, models/loss.py:26-70, models/model06.py:38-72            #
# --------------------------------------------------------------------------- #
def logmeanexp(log_w: torch.Tensor, axis: int) -> torch.Tensor:
    """utils/utils.py:9-11 (no stop_gradient on the max: the gradient is softmax over ``axis``)."""
    mx = torch.amax(log_w, dim=axis)                                             # :10
    return torch.log(torch.mean(torch.exp(log_w - mx.unsqueeze(axis)), dim=axis)) + mx  # :11


def iwae_loss(lpxz_elem: torch.Tensor, lpz: torch.Tensor, lqzx: torch.Tensor, x_shape, pxz_axes=(-1, -2, -3), beta=1.0):
    """models/loss.py:26-55 from the point where the three ``log_prob`` tensors exist.

    ``lpxz_elem`` = ``pxz.log_prob(x)``; ``lpz``/``lqzx`` already summed over their axes (``[S,B]``).
    """
    lpxz = torch.sum(lpxz_elem, dim=tuple(pxz_axes))                             # :32
    log_w = lpxz + beta * (lpz - lqzx)                                           # :34
    iwae_elbo = torch.mean(logmeanexp(log_w, axis=0), dim=-1)                    # :37
    n_dims = float(math.prod(list(x_shape)[1:]))                                 # :42
    bpd = -iwae_elbo / (math.log(2.0) * n_dims)                                  # :43
    kl = -torch.mean(lpz - lqzx, dim=0)                                          # :46
    return -iwae_elbo, {"iwae_elbo": iwae_elbo, "bpd": bpd, "lpxz": lpxz, "lqzx": lqzx, "lpz": lpz, "kl": kl}


def elbo_loss(lpxz_elem: torch.Tensor, lpz: torch.Tensor, lqzx: torch.Tensor, pxz_axes=(-1, -2, -3)):
    """models/loss.py:58-70."""
    lpxz = torch.sum(lpxz_elem, dim=tuple(pxz_axes))                             # :64
    log_w = lpxz + (lpz - lqzx)                                                  # :66
    elbo = torch.mean(torch.mean(log_w, dim=0), dim=-1)                          # :68
    return -elbo, {"loss": -elbo, "lpxz": lpxz}


def model06_loss(lpxz_elem, lpz2, lqz2z1, lpz1z2, lqz1x, x_shape, pxz_axes=(-1, -2, -3)):
    """models/model06.py:38-72 from the point where the ``log_prob`` sums exist."""
    lpxz = torch.sum(lpxz_elem, dim=tuple(pxz_axes))                             # :45
    log_w = lpxz + (lpz2 - lqz2z1) + (lpz1z2 - lqz1x)                            # :47
    iwae_elbo = torch.mean(logmeanexp(log_w, axis=0), dim=-1)                    # :50
    n_dims = float(math.prod(list(x_shape)[-len(pxz_axes):]))                    # :54
    bpd = -iwae_elbo / (math.log(2.0) * n_dims)                                  # :55
    kl1 = -torch.mean(lpz1z2 - lqz1x, dim=0)                                     # :58
    kl2 = -torch.mean(lpz2 - lqz2z1, dim=0)                                      # :59
    return -iwae_elbo, {
        "iwae_elbo": iwae_elbo, "bpd": bpd, "lpxz": lpxz, "lqz1x": lqz1x, "lqz2z1": lqz2z1,
        "lpz2": lpz2, "lpz1z2": lpz1z2, "kl1": kl1, "kl2": kl2,
    }


# --------------------------------------------------------------------------- #
# data contract + the quantiser (new-build definition)               #
# --------------------------------------------------------------------------- #
def normalize_u8(x_u8: torch.Tensor, dtype=torch.float32) -> torch.Tensor:
    """utils/data.py:15-16: ``cast(img, float32) / 255.``"""
    return x_u8.to(dtype) / 255.0


def quantise(x01: torch.Tensor) -> torch.Tensor:
    """``uint8(rint(255 * clip(x01, 0, 1)))`` -- the reference never quantises
    (utils/mdl_openai.py:184); this is the build's definition, evaluated in float64."""
    return torch.round(255.0 * torch.clamp(x01.double(), 0.0, 1.0)).to(torch.uint8)
