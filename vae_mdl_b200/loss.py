"""Mirror of models/loss.py (``iwae_loss`` :26-55, ``elbo_loss`` :58-70) and of model06's ``loss_fn``
(models/model06.py:38-72), plus the fused observation-model step used by the benchmark.

Like the reference (models/loss.py:13-23) importing this module gives every ``torch.distributions.Distribution`` a
settable ``axes`` property, so ``pz`` / ``qzx`` can be plain ``torch.distributions.Normal`` objects.
"""
from __future__ import annotations

import math

import torch
import torch.distributions as td

from . import functional as F
from .utils import logmeanexp

__all__ = ["iwae_loss", "elbo_loss", "loss_fn", "modl_iwae_step", "dlogistic_iwae_step"]


def _get_axes(self):
    return self._axes


def _set_axes(self, axes):
    self._axes = axes


td.Distribution.axes = property(_get_axes, _set_axes)  # models/loss.py:13-23


def _axes(d):
    return tuple(d.axes)


def _lpxz(pxz, x):
    """``reduce_sum(pxz.log_prob(x), axis=pxz.axes)`` (models/loss.py:32).  The observation models of this package
    compute the per-image sum inside the kernel when the axes are the image axes."""
    axes = [a for a in pxz.axes]
    if hasattr(pxz, "log_likelihood") and sorted(axes) == [-3, -2, -1]:
        # float64-accumulated sums: log_w ~ -2e4 has a float32 ulp of 2e-3, which a float32 lpxz would feed straight
        # into the softmax over importance samples.  The metric dict still reports float32 (reference dtype).
        return pxz.log_likelihood(x, dtype=torch.float64)
    return torch.sum(pxz.log_prob(x), dim=tuple(axes))


def _is_std_const(dist):
    """``Normal(0.0, 1.0)`` built from Python numbers (the reference's prior, models/model05.py:108): 0-dim CPU tensors."""
    loc, scale = dist.loc, dist.scale
    return (not loc.is_cuda and not scale.is_cuda and loc.numel() == 1 and scale.numel() == 1
            and not loc.requires_grad and not scale.requires_grad and float(loc) == 0.0 and float(scale) == 1.0)


def _normal_ok(dist, axes, z):
    """A plain Normal density (not a subclass / transformed / Independent wrapper) reduced over the last axis of a CUDA
    ``[S,B,D]`` float sample; parameters not wider than the sample and either on z's device or the constant (0, 1)."""
    if type(dist) is not td.Normal or list(axes) != [-1] or z is None or z.dim() != 3:
        return False
    if not z.is_cuda or not z.is_floating_point():
        return False
    if _is_std_const(dist):
        return True
    for t in (dist.loc, dist.scale):
        if not t.is_floating_point() or t.device != z.device:
            return False
        if t.dim() > 3 or (t.dim() == 3 and t.shape[0] not in (1, z.shape[0])):
            return False
    return True


def _term(z, dist, weight):
    """One Normal density of ``log_w`` for ``F.fused_iwae_loss``; the constant (0, 1) prior is the parameter-free term."""
    if _is_std_const(dist):
        return (z, None, None, weight)
    return (z, dist.loc, dist.scale, weight)


def _obs_ok(pxz, axes, x, S, B):
    """One of this package's observation models scored over the image axes, parameters ``[S,B,H,W,...]``, x ``[B,H,W,3]``."""
    if not hasattr(pxz, "_iwae_spec") or sorted(axes) != [-3, -2, -1]:
        return False
    _, _, p0, p1 = pxz._iwae_spec()
    if not (p0.dim() == 5 and x.dim() == 4 and p0.shape[0] == S and p0.shape[1] == B and x.shape[0] == B):
        return False
    return p0.is_cuda and x.is_cuda and x.device == p0.device and (p1 is None or p1.device == p0.device)


def _fusable(x, z, pz, qzx, pxz):
    if not (_normal_ok(pz, _axes(pz), z) and _normal_ok(qzx, _axes(qzx), z)
            and _obs_ok(pxz, pxz.axes, x, z.shape[0], z.shape[1])):
        return False
    return z.device == x.device


def iwae_loss(x, z, pz, qzx, pxz, beta=1.0):
    """models/loss.py:26-55; returns ``(-iwae_elbo, metrics)`` with the reference's metric keys.

    With this package's observation models and Normal latents the whole objective after the networks runs fused:
    latent terms (1 launch) -> observation-model forward -> finish (per-image sums, log-mean-exp, batch mean), and two
    launches backward; anything else takes the generic op-by-op route below."""
    if _fusable(x, z, pz, qzx, pxz):
        kind, meta, p0, p1 = pxz._iwae_spec()
        loss, lpxz, sums, _ = F.fused_iwae_loss(kind, meta, x, p0, p1, [_term(z, pz, beta),               # :34
                                                                        _term(z, qzx, -beta)])
        neg_elbo = loss.detach()
        lpz, lqzx = sums[0], sums[1]
        n_dims = float(math.prod(x.shape[1:]))                              # :42
        return loss, {"iwae_elbo": -neg_elbo, "bpd": neg_elbo * (1.0 / (math.log(2.0) * n_dims)), "lpxz": lpxz,  # :43
                      "lqzx": lqzx, "lpz": lpz, "kl": torch.mean(lqzx - lpz, dim=0)}                                # :46
    lpz = torch.sum(pz.log_prob(z), dim=_axes(pz))                      # :28
    lqzx = torch.sum(qzx.log_prob(z), dim=_axes(qzx))                   # :30
    lpxz = _lpxz(pxz, x)                                                # :32
    log_w = lpxz + beta * (lpz - lqzx)                                  # :34
    iwae_elbo = torch.mean(logmeanexp(log_w, axis=0), dim=-1)           # :37
    n_dims = float(math.prod(x.shape[1:]))                              # :42 (reference semantics kept)
    bpd = -iwae_elbo / (math.log(2.0) * n_dims)                         # :43
    kl = -torch.mean(lpz - lqzx, dim=0)                                 # :46
    return -iwae_elbo, {"iwae_elbo": iwae_elbo, "bpd": bpd, "lpxz": lpxz.float(), "lqzx": lqzx, "lpz": lpz, "kl": kl}


def elbo_loss(x, z, pz, qzx, pxz):
    """models/loss.py:58-70."""
    lpz = torch.sum(pz.log_prob(z), dim=_axes(pz))
    lqzx = torch.sum(qzx.log_prob(z), dim=_axes(qzx))
    lpxz = _lpxz(pxz, x)
    log_w = lpxz + (lpz - lqzx)
    elbo = torch.mean(torch.mean(log_w, dim=0), dim=-1).float()
    return -elbo, {"loss": -elbo, "lpxz": lpxz.float()}


def loss_fn(x, pz, qz1x, qz2z1, pz1z2, pxz1):
    """model06's loss (models/model06.py:38-72); arguments are ``DistributionTuple``s except ``pz``.  Runs fused (latent
    terms -> observation model -> finish) when the four latent densities are Normals over the last axis."""
    z1, z2 = qz1x.z, qz2z1.z
    if (z1 is not None and z2 is not None and _normal_ok(qz2z1.dist, qz2z1.axes, z2) and _normal_ok(qz1x.dist, qz1x.axes, z1)
            and _normal_ok(pz, _axes(pz), z2) and _normal_ok(pz1z2.dist, qz1x.axes, z1)
            and _obs_ok(pxz1.dist, pxz1.axes, x, z1.shape[0], z1.shape[1]) and z1.device == x.device == z2.device):
        kind, meta, p0, p1 = pxz1.dist._iwae_spec()
        terms = [_term(z2, pz, 1.0), _term(z2, qz2z1.dist, -1.0),                                          # :47
                 _term(z1, pz1z2.dist, 1.0), _term(z1, qz1x.dist, -1.0)]
        loss, lpxz, sums, _ = F.fused_iwae_loss(kind, meta, x, p0, p1, terms)
        iwae_elbo = -loss.detach()
        lpz2, lqz2z1, lpz1z2, lqz1x = sums[0], sums[1], sums[2], sums[3]
        n_dims = float(math.prod(x.shape[-len(pxz1.axes):]))            # :54
        return loss, {"iwae_elbo": iwae_elbo, "bpd": -iwae_elbo / (math.log(2.0) * n_dims), "lpxz": lpxz, "lqz1x": lqz1x,
                      "lqz2z1": lqz2z1, "lpz2": lpz2, "lpz1z2": lpz1z2, "kl1": -torch.mean(lpz1z2 - lqz1x, dim=0),
                      "kl2": -torch.mean(lpz2 - lqz2z1, dim=0)}
    lqz2z1 = torch.sum(qz2z1.dist.log_prob(qz2z1.z), dim=tuple(qz2z1.axes))
    lqz1x = torch.sum(qz1x.dist.log_prob(qz1x.z), dim=tuple(qz1x.axes))
    lpz2 = torch.sum(pz.log_prob(qz2z1.z), dim=_axes(pz))
    lpz1z2 = torch.sum(pz1z2.dist.log_prob(qz1x.z), dim=tuple(qz1x.axes))
    dist = pxz1.dist
    if hasattr(dist, "log_likelihood") and sorted(pxz1.axes) == [-3, -2, -1]:
        lpxz = dist.log_likelihood(x, dtype=torch.float64)
    else:
        lpxz = torch.sum(dist.log_prob(x), dim=tuple(pxz1.axes))
    log_w = lpxz + (lpz2 - lqz2z1) + (lpz1z2 - lqz1x)                    # :47
    iwae_elbo = torch.mean(logmeanexp(log_w, axis=0), dim=-1)           # :50
    n_dims = float(math.prod(x.shape[-len(pxz1.axes):]))                # :54
    bpd = -iwae_elbo / (math.log(2.0) * n_dims)
    kl1 = -torch.mean(lpz1z2 - lqz1x, dim=0)
    kl2 = -torch.mean(lpz2 - lqz2z1, dim=0)
    return -iwae_elbo, {"iwae_elbo": iwae_elbo, "bpd": bpd, "lpxz": lpxz.float(), "lqz1x": lqz1x, "lqz2z1": lqz2z1,
                        "lpz2": lpz2, "lpz1z2": lpz1z2, "kl1": kl1, "kl2": kl2}


def modl_iwae_step(params: torch.Tensor, x: torch.Tensor, extra: torch.Tensor = None, need_grad: bool = True,
                   b_total: int = 0):
    """The whole observation-model side of one IWAE step, no autograd graph: MoDL forward (tile partial sums, float64) ->
    finish (per-image sums, log-mean-exp, elbo, softmax weights) -> MoDL gradient; ONE cooperative kernel launch for the
    training shapes of models/model05.py (``vaemdl_modl_iwae_step``), three launches otherwise.

    ``params [S,B,H,W,10M]``, ``x [B,H,W,3]`` (uint8 or float in [0,1]), ``extra = beta*(lpz-lqzx) [S,B]`` or None.
    ``b_total``: whole-batch size when ``params`` is one rank's batch shard (the returned loss is then this rank's
    additive share of the global loss).
    Returns ``(loss=-elbo [1], lpxz [S,B] float64, dparams or None)`` -- the numbers ``iwae_loss`` + ``backward`` give.
    """
    with torch.no_grad():
        if need_grad:  # one call: a single cooperative launch for training shapes, forward -> finish -> gradient otherwise
            lpxz, _, _, elbo, _, dparams, _ = F.modl_iwae_step(params, x, extra, b_total)
        else:
            lpxz, _, _, elbo, _ = F.modl_iwae_forward(params, x, extra, b_total)
            dparams = None
    return -elbo, lpxz, dparams


def dlogistic_iwae_step(loc: torch.Tensor, logscale: torch.Tensor, x: torch.Tensor, extra: torch.Tensor = None,
                        low=-1.0, high=1.0, levels=256.0, need_grad: bool = True, b_total: int = 0):
    """The plain discretized-logistic counterpart of ``modl_iwae_step`` (models 03/04/06): forward -> finish -> gradient,
    one cooperative launch for image shapes (``vaemdl_dlogistic_iwae_step``), three launches otherwise.  ``loc``/``logscale [S,B,H,W,3]`` (e.g. the two halves of the ``[..,6]`` conv output,
    models/model03.py:88-91), ``x [B,H,W,3]``.  Returns ``(loss=-elbo [1], lpxz [S,B] float64, dloc, dlogscale)``."""
    with torch.no_grad():
        if need_grad:  # one call: a single cooperative launch for the image shapes of models 03/04/06
            lpxz, _, _, elbo, _, dloc, dls, _ = F.dlogistic_iwae_step(loc, logscale, x, extra, low, high, levels, b_total)
        else:
            lpxz, _, _, elbo, _ = F.dlogistic_iwae_forward(loc, logscale, x, extra, low, high, levels, b_total)
            dloc = dls = None
    return -elbo, lpxz, dloc, dls
