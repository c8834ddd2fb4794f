"""Tensor-level entry points: shape handling + ``torch.autograd.Function`` wrappers around the C ABI.

Everything here enqueues hand-written sm_100a kernels from ``libvaemdl_b200.so`` on the current CUDA
stream; PyTorch only owns the buffers.  CPU tensors are rejected (``_abi.require_cuda``).
"""
from __future__ import annotations

import ctypes
import math
from typing import Optional, Tuple

import torch

from . import _abi
from ._abi import check, dense_f32, dense_param, lib, ptr, stream_ptr

__all__ = [
    "modl_log_prob",
    "modl_log_likelihood",
    "modl_backward",
    "modl_iwae_forward",
    "dlogistic_log_prob",
    "dlogistic_log_likelihood",
    "dlogistic_iwae_forward",
    "dlogistic_backward",
    "logmeanexp",
    "iwae_tail",
    "modl_sample",
    "dlogistic_sample",
    "latent_terms",
    "latent_terms_backward",
    "fused_iwae_loss",
]


# --------------------------------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------------------------------
def _prep_x(x: torch.Tensor, img_shape: Tuple[int, ...], what: str):
    """Returns (x dense, x_dtype enum, x_batch).  x: [x_batch, *img_shape] or [*img_shape]; uint8 or float."""
    _abi.require_cuda(x, what)
    nd = len(img_shape)
    if tuple(x.shape[-nd:]) != tuple(img_shape):
        raise ValueError(f"{what}: trailing dims {tuple(x.shape[-nd:])} do not match the parameter image shape {img_shape}")
    lead = x.shape[:-nd]
    x_batch = int(math.prod(lead)) if len(lead) else 1
    if x.dtype == torch.uint8:
        return x.contiguous(), _abi.X_U8, x_batch
    return dense_f32(x, what), _abi.X_F32, x_batch


def _check_batch(n_img: int, x_batch: int, what: str):
    if n_img % x_batch:
        raise ValueError(f"{what}: {n_img} parameter images cannot be scored against {x_batch} observed images "
                         "(the leading dims of the parameters must end with the batch dim of x)")


def _bins(plain):
    """``plain``: False (x-conditioned classes), True (utils/mdl_plain.py with its default low=-1, high=1, levels=256) or
    the ``(low, high, levels)`` the un-conditioned class was built with (utils/mdl_plain.py:18)."""
    if isinstance(plain, (tuple, list)):
        low, high, levels = (float(v) for v in plain)
        return low, high, levels
    return -1.0, 1.0, 256.0


class _ModlFn(torch.autograd.Function):
    """Forward: per-pixel log-prob or per-image log-likelihood (float32 or float64 sums).  Backward: one fused kernel."""

    @staticmethod
    def forward(ctx, params, x, x_range, edge_mode, mode, plain=False):
        # mode: "pixel" -> [..., H, W] ; "image" -> [...] float32 ; "image64" -> [...] float64
        # plain: utils/mdl_plain.py (means chained on the means) instead of utils/mdl.py (chained on the observed x)
        p, bf16 = dense_param(params, "parameters")
        if bf16 and plain:
            raise ValueError("bfloat16 parameters are implemented for the x-conditioned mixture classes only")
        H, W, C10 = p.shape[-3], p.shape[-2], p.shape[-1]
        M = C10 // 10
        if C10 != 10 * M or M < 1:
            raise ValueError(f"last parameter dim must be 10*n_mix, got {C10}")
        lead = tuple(p.shape[:-3])
        n_img = int(math.prod(lead)) if lead else 1
        xd, x_dtype, x_batch = _prep_x(x, (H, W, 3), "x")
        _check_batch(n_img, x_batch, "log_prob")
        L = lib()
        lp = ll = ll64 = ws = None
        ws_bytes = 0
        if mode == "pixel":
            lp = torch.empty(lead + (H, W), device=p.device, dtype=torch.float32)
        else:
            if mode == "image64":
                ll64 = torch.empty(lead, device=p.device, dtype=torch.float64)
            else:
                ll = torch.empty(lead, device=p.device, dtype=torch.float32)
            ws_bytes = L.vaemdl_modl_workspace_bytes(n_img, H, W)
            ws = torch.empty((ws_bytes + 7) // 8, device=p.device, dtype=torch.float64)
        with _abi.on_device(p.device):
            if plain:
                if x_range != _abi.RANGE_UNIT or edge_mode != _abi.EDGE_MDL:
                    raise ValueError("the plain pixel mixture takes x in [0,1] and the <= -1 / >= 1 edge tests")
                check(L.vaemdl_modl_plain_fwd(ptr(p), ptr(xd), x_dtype, n_img, x_batch, H, W, M, *_bins(plain), ptr(lp), ptr(ll),
                                              ptr(ll64), ptr(ws), ws_bytes, stream_ptr(p.device)), "vaemdl_modl_plain_fwd")
            else:
                fn = L.vaemdl_modl_fwd_bf16 if bf16 else L.vaemdl_modl_fwd
                check(fn(ptr(p), ptr(xd), x_dtype, x_range, edge_mode, n_img, x_batch, H, W, M,
                         ptr(lp), ptr(ll), ptr(ll64), ptr(ws), ws_bytes, stream_ptr(p.device)), "vaemdl_modl_fwd")
        ctx.save_for_backward(p, xd)
        ctx.meta = (x_dtype, x_range, edge_mode, n_img, x_batch, H, W, M, mode, plain, bf16)
        return lp if mode == "pixel" else (ll64 if mode == "image64" else ll)

    @staticmethod
    def backward(ctx, g):
        p, xd = ctx.saved_tensors
        x_dtype, x_range, edge_mode, n_img, x_batch, H, W, M, mode, plain, bf16 = ctx.meta
        if g is None:
            return None, None, None, None, None, None
        g = dense_f32(g, "upstream gradient")
        g_pixel, g_image = (g, None) if mode == "pixel" else (None, g)
        dp = torch.empty_like(p)
        with _abi.on_device(p.device):
            if plain:
                check(lib().vaemdl_modl_plain_bwd(ptr(p), ptr(xd), x_dtype, n_img, x_batch, H, W, M, *_bins(plain),
                                                  ptr(g_image), ptr(g_pixel), ptr(dp), stream_ptr(p.device)),
                      "vaemdl_modl_plain_bwd")
            else:
                fn = lib().vaemdl_modl_bwd_bf16 if bf16 else lib().vaemdl_modl_bwd  # dp has the dtype of p
                check(fn(ptr(p), ptr(xd), x_dtype, x_range, edge_mode, n_img, x_batch, H, W, M,
                         ptr(g_image), ptr(g_pixel), ptr(dp), stream_ptr(p.device)), "vaemdl_modl_bwd")
        return dp, None, None, None, None, None


def modl_backward(params: torch.Tensor, x: torch.Tensor, g_image: Optional[torch.Tensor] = None,
                  g_pixel: Optional[torch.Tensor] = None, x_range: int = _abi.RANGE_UNIT,
                  edge_mode: int = _abi.EDGE_MDL, plain: bool = False,
                  pix_stats: Optional[torch.Tensor] = None) -> torch.Tensor:
    """The gradient kernel on its own: d/dparams of sum(g_image * ll_image) + sum(g_pixel * lp_pixel).
    bfloat16 parameters give a bfloat16 gradient (float32 arithmetic, one rounding at the store)."""
    p, bf16 = dense_param(params, "parameters")
    if bf16 and plain:
        raise ValueError("bfloat16 parameters are implemented for the x-conditioned mixture classes only")
    H, W, C10 = p.shape[-3:]
    lead = tuple(p.shape[:-3])
    n_img = int(math.prod(lead)) if lead else 1
    xd, x_dtype, x_batch = _prep_x(x, (H, W, 3), "x")
    _check_batch(n_img, x_batch, "backward")
    gi = dense_f32(g_image, "g_image") if g_image is not None else None
    gp = dense_f32(g_pixel, "g_pixel") if g_pixel is not None else None
    dp = torch.empty_like(p)
    with _abi.on_device(p.device):
        if plain:
            check(lib().vaemdl_modl_plain_bwd(ptr(p), ptr(xd), x_dtype, n_img, x_batch, H, W, C10 // 10, *_bins(plain), ptr(gi),
                                              ptr(gp), ptr(dp), stream_ptr(p.device)), "vaemdl_modl_plain_bwd")
        elif pix_stats is not None:
            # the per-pixel mixture sums of the forward call on the same parameters: one-pass gradient kernel (bfloat16
            # parameters: the tile then stays bfloat16 in shared memory, two slots per warp, gradient rounded once)
            fn = lib().vaemdl_modl_bwd_stats_bf16 if bf16 else lib().vaemdl_modl_bwd_stats
            check(fn(ptr(p), ptr(xd), x_dtype, x_range, edge_mode, n_img, x_batch, H, W, C10 // 10,
                     ptr(gi), ptr(gp), ptr(pix_stats), ptr(dp), stream_ptr(p.device)), "vaemdl_modl_bwd_stats")
        else:
            fn = lib().vaemdl_modl_bwd_bf16 if bf16 else lib().vaemdl_modl_bwd
            check(fn(ptr(p), ptr(xd), x_dtype, x_range, edge_mode, n_img, x_batch, H, W, C10 // 10,
                     ptr(gi), ptr(gp), ptr(dp), stream_ptr(p.device)), "vaemdl_modl_bwd")
    return dp


def modl_iwae_forward(params: torch.Tensor, x: torch.Tensor, extra: Optional[torch.Tensor] = None, b_total: int = 0,
                      x_range: int = _abi.RANGE_UNIT, edge_mode: int = _abi.EDGE_MDL, plain: bool = False,
                      want_stats: bool = False):
    """MoDL forward fused with the IWAE tail in TWO launches (``vaemdl_modl_iwae_fwd``): ``params [S,B,H,W,10M]``,
    ``x [B,H,W,3]``, ``extra = beta*(lpz-lqzx) [S,B]`` or None.  Returns ``(lpxz float64 [S,B], log_w, lme_b [B],
    elbo [1], g_ll [S,B])`` with ``g_ll = d(-elbo)/d lpxz`` (models/loss.py:32-37).  Not recorded by autograd."""
    p, bf16 = dense_param(params, "parameters")
    if bf16 and plain:
        raise ValueError("bfloat16 parameters are implemented for the x-conditioned mixture classes only")
    if p.dim() != 5:
        raise ValueError("parameters must be [S, B, H, W, 10*n_mix]")
    S, B, H, W, C10 = p.shape
    M = C10 // 10
    if C10 != 10 * M or M < 1:
        raise ValueError(f"last parameter dim must be 10*n_mix, got {C10}")
    xd, x_dtype, x_batch = _prep_x(x, (H, W, 3), "x")
    if x_batch not in (1, B):
        raise ValueError(f"x must hold {B} images (or one), got {x_batch}")
    ex = dense_f32(extra, "extra").reshape(S, B) if extra is not None else None
    L = lib()
    dev = p.device
    ll64 = torch.empty((S, B), device=dev, dtype=torch.float64)
    log_w = torch.empty((S, B), device=dev, dtype=torch.float32)
    lme_b = torch.empty(B, device=dev, dtype=torch.float32)
    elbo = torch.empty(1, device=dev, dtype=torch.float32)
    g_ll = torch.empty((S, B), device=dev, dtype=torch.float32)
    ws_bytes = L.vaemdl_modl_workspace_bytes(S * B, H, W)
    ws = torch.empty((ws_bytes + 7) // 8, device=dev, dtype=torch.float64)
    with _abi.on_device(dev):
        if plain:
            check(L.vaemdl_modl_plain_iwae_fwd(ptr(p), ptr(xd), x_dtype, S, B, int(b_total), x_batch, H, W, M, *_bins(plain),
                                               ptr(ex), None,
                                               ptr(ll64), ptr(log_w), ptr(lme_b), ptr(elbo), ptr(g_ll), ptr(ws), ws_bytes,
                                               stream_ptr(dev)), "vaemdl_modl_plain_iwae_fwd")
        else:
            if want_stats:
                # leaves every pixel's (mixture sum, logit normaliser) for modl_backward(pix_stats=...): one-pass gradient
                stats = torch.empty((S, B, H, W, 2), device=dev, dtype=torch.float32)
                fn = L.vaemdl_modl_iwae_fwd_stats_bf16 if bf16 else L.vaemdl_modl_iwae_fwd_stats
                check(fn(ptr(p), ptr(xd), x_dtype, x_range, edge_mode, S, B, int(b_total), x_batch,
                         H, W, M, ptr(ex), None, ptr(ll64), ptr(log_w), ptr(lme_b), ptr(elbo),
                         ptr(g_ll), ptr(stats), ptr(ws), ws_bytes, stream_ptr(dev)), "vaemdl_modl_iwae_fwd_stats")
                return ll64, log_w, lme_b, elbo, g_ll, stats
            fn = L.vaemdl_modl_iwae_fwd_bf16 if bf16 else L.vaemdl_modl_iwae_fwd
            check(fn(ptr(p), ptr(xd), x_dtype, x_range, edge_mode, S, B, int(b_total), x_batch, H, W, M,
                     ptr(ex), None, ptr(ll64), ptr(log_w), ptr(lme_b), ptr(elbo), ptr(g_ll), ptr(ws),
                     ws_bytes, stream_ptr(dev)), "vaemdl_modl_iwae_fwd")
    if want_stats:
        return ll64, log_w, lme_b, elbo, g_ll, None
    return ll64, log_w, lme_b, elbo, g_ll


def modl_iwae_step(params: torch.Tensor, x: torch.Tensor, extra: Optional[torch.Tensor] = None, b_total: int = 0,
                   x_range: int = _abi.RANGE_UNIT, edge_mode: int = _abi.EDGE_MDL, plain: bool = False):
    """Forward + IWAE finish + parameter gradient in ONE call (``vaemdl_modl_iwae_step``): one cooperative kernel launch
    for the training shapes of models/model05.py, three launches otherwise.  Arguments as ``modl_iwae_forward``.
    Returns ``(lpxz float64 [S,B], log_w, lme_b [B], elbo [1], g_ll [S,B], dparams, launches)``."""
    if params.dtype == torch.bfloat16:  # bfloat16 parameters: forward + finish, then the gradient kernel (3 launches)
        # n_mix 10 / 20 / 30: the forward kernel leaves the per-pixel sums and the gradient kernel keeps the tile in bfloat16
        want = (not plain) and params.shape[-1] in (60, 80, 100, 120, 140, 160, 180, 200, 240, 280, 300, 320, 360, 400, 480, 500, 560, 600, 640)
        out = modl_iwae_forward(params, x, extra, b_total, x_range, edge_mode, plain, want_stats=want)
        ll64, log_w, lme_b, elbo, g_ll = out[:5]
        dp = modl_backward(params, x, g_image=g_ll, x_range=x_range, edge_mode=edge_mode, pix_stats=out[5] if want else None)
        return ll64, log_w, lme_b, elbo, g_ll, dp, 3
    p = dense_f32(params, "parameters")
    if p.dim() != 5:
        raise ValueError("parameters must be [S, B, H, W, 10*n_mix]")
    S, B, H, W, C10 = p.shape
    M = C10 // 10
    if C10 != 10 * M or M < 1:
        raise ValueError(f"last parameter dim must be 10*n_mix, got {C10}")
    xd, x_dtype, x_batch = _prep_x(x, (H, W, 3), "x")
    if x_batch not in (1, B):
        raise ValueError(f"x must hold {B} images (or one), got {x_batch}")
    ex = dense_f32(extra, "extra").reshape(S, B) if extra is not None else None
    L = lib()
    dev = p.device
    ll64 = torch.empty((S, B), device=dev, dtype=torch.float64)
    log_w = torch.empty((S, B), device=dev, dtype=torch.float32)
    lme_b = torch.empty(B, device=dev, dtype=torch.float32)
    elbo = torch.empty(1, device=dev, dtype=torch.float32)
    g_ll = torch.empty((S, B), device=dev, dtype=torch.float32)
    dp = torch.empty_like(p)
    ws_bytes = L.vaemdl_modl_step_workspace_bytes(S * B, H, W)   # incl. the per-pixel sums the two passes share
    ws = torch.empty((ws_bytes + 7) // 8, device=dev, dtype=torch.float64)
    n_launch = ctypes.c_int(0)
    with _abi.on_device(dev):
        if plain:
            check(L.vaemdl_modl_plain_iwae_step(ptr(p), ptr(xd), x_dtype, S, B, int(b_total), x_batch, H, W, M,
                                                *_bins(plain), ptr(ex),
                                                None, ptr(ll64), ptr(log_w), ptr(lme_b), ptr(elbo), ptr(g_ll), ptr(dp),
                                                ptr(ws), ws_bytes, stream_ptr(dev), ctypes.byref(n_launch)),
                  "vaemdl_modl_plain_iwae_step")
        else:
            check(L.vaemdl_modl_iwae_step(ptr(p), ptr(xd), x_dtype, x_range, edge_mode, S, B, int(b_total), x_batch, H, W,
                                          M, ptr(ex), None, ptr(ll64), ptr(log_w), ptr(lme_b), ptr(elbo), ptr(g_ll), ptr(dp),
                                          ptr(ws), ws_bytes, stream_ptr(dev), ctypes.byref(n_launch)),
                  "vaemdl_modl_iwae_step")
    return ll64, log_w, lme_b, elbo, g_ll, dp, n_launch.value


def modl_log_prob(params: torch.Tensor, x: torch.Tensor, x_range: int = _abi.RANGE_UNIT,
                  edge_mode: int = _abi.EDGE_MDL, plain: bool = False) -> torch.Tensor:
    """Per-pixel MoDL log-prob ``[..., H, W]`` (utils/mdl.py:56-92 without the trailing ``expand_dims``;
    ``plain=True``: utils/mdl_plain.py:36-66)."""
    return _ModlFn.apply(params, x, x_range, edge_mode, "pixel", plain)


def modl_log_likelihood(params: torch.Tensor, x: torch.Tensor, x_range: int = _abi.RANGE_UNIT,
                        edge_mode: int = _abi.EDGE_MDL, dtype: torch.dtype = torch.float32,
                        plain: bool = False) -> torch.Tensor:
    """Per-image MoDL log-likelihood ``[...]`` = ``reduce_sum(log_prob(x), [-1,-2,-3])`` (models/loss.py:32),
    computed without ever writing the per-pixel tensor.  ``dtype=torch.float64`` returns the float64-accumulated sums
    (same float32 per-pixel values; no float32 rounding of the ~-2e4 totals)."""
    return _ModlFn.apply(params, x, x_range, edge_mode, "image64" if dtype == torch.float64 else "image", plain)


# --------------------------------------------------------------------------------------------------
# plain discretized logistic
# --------------------------------------------------------------------------------------------------
def _dl_layout(loc: torch.Tensor, logscale: torch.Tensor):
    """Detects the un-split ``[..., 2C]`` conv output (models/model03.py:88-91): ``loc`` and ``logscale`` are the two
    halves of one contiguous tensor.  Returns (loc, logscale, C, ld) with both tensors addressable as
    ``base[(e // C) * ld + e % C]``."""
    C = loc.shape[-1]
    if (loc.dtype == torch.float32 and logscale.dtype == torch.float32 and loc.shape == logscale.shape and loc.dim() >= 1
            and loc.stride() == logscale.stride() and loc.stride(-1) == 1
            and logscale.data_ptr() == loc.data_ptr() + 4 * C and loc.data_ptr() % 16 == 0):
        # candidate split view: check all leading strides describe a dense [..., 2C] parent
        expect = 2 * C
        okay = True
        for size, stride in zip(reversed(loc.shape[:-1]), reversed(loc.stride()[:-1])):
            if size != 1 and stride != expect:
                okay = False
                break
            expect *= size
        if okay:
            return loc, logscale, C, 2 * C
    return dense_f32(loc, "loc"), dense_f32(logscale, "logscale"), C, C


class _DlFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, loc, logscale, x, low, high, levels, n_event_dims, mode):
        if loc.shape != logscale.shape:
            loc, logscale = torch.broadcast_tensors(loc, logscale)
        _abi.require_cuda(loc, "loc")
        locd, lsd, C, ld = _dl_layout(loc, logscale)
        shape = tuple(loc.shape)
        ev = shape[len(shape) - n_event_dims:]
        lead = shape[:len(shape) - n_event_dims]
        D = int(math.prod(ev))
        n_img = int(math.prod(lead)) if lead else 1
        xd, x_dtype, x_batch = _prep_x(x, ev, "x")
        _check_batch(n_img, x_batch, "log_prob")
        L = lib()
        lp = ll = ll64 = ws = None
        ws_bytes = 0
        if mode == "elem":
            lp = torch.empty(shape, device=loc.device, dtype=torch.float32)
        else:
            if mode == "image64":
                ll64 = torch.empty(lead, device=loc.device, dtype=torch.float64)
            else:
                ll = torch.empty(lead, device=loc.device, dtype=torch.float32)
            ws_bytes = L.vaemdl_dlogistic_workspace_bytes(n_img, D)
            ws = torch.empty((ws_bytes + 7) // 8, device=loc.device, dtype=torch.float64)
        with _abi.on_device(loc.device):
            check(L.vaemdl_dlogistic_fwd(ptr(locd), ptr(lsd), C, ld, ptr(xd), x_dtype, n_img, x_batch, D,
                                         float(low), float(high), float(levels), ptr(lp), ptr(ll), ptr(ll64), ptr(ws),
                                         ws_bytes, stream_ptr(loc.device)), "vaemdl_dlogistic_fwd")
        ctx.save_for_backward(locd, lsd, xd)
        ctx.meta = (C, ld, x_dtype, n_img, x_batch, D, float(low), float(high), float(levels), shape, mode)
        return lp if mode == "elem" else (ll64 if mode == "image64" else ll)

    @staticmethod
    def backward(ctx, g):
        locd, lsd, xd = ctx.saved_tensors
        C, ld, x_dtype, n_img, x_batch, D, low, high, levels, shape, mode = ctx.meta
        if g is None:
            return (None,) * 8
        g = dense_f32(g, "upstream gradient")
        g_elem, g_image = (g, None) if mode == "elem" else (None, g)
        # gradients of the two halves of an un-split [..,2C] tensor are written into one [..,2C] buffer
        if ld == 2 * C:
            both = torch.empty(shape[:-1] + (2 * C,), device=locd.device, dtype=torch.float32)
            dloc, dls, ld_out = both[..., :C], both[..., C:], 2 * C
            p_loc, p_ls = both.data_ptr(), both.data_ptr() + 4 * C
        else:
            dloc = torch.empty(shape, device=locd.device, dtype=torch.float32)
            dls = torch.empty(shape, device=locd.device, dtype=torch.float32)
            ld_out = C
            p_loc, p_ls = dloc.data_ptr(), dls.data_ptr()
        import ctypes
        with _abi.on_device(locd.device):
            check(lib().vaemdl_dlogistic_bwd(ptr(locd), ptr(lsd), C, ld, ptr(xd), x_dtype, n_img, x_batch, D, low, high,
                                             levels, ptr(g_image), ptr(g_elem), ctypes.c_void_p(p_loc),
                                             ctypes.c_void_p(p_ls), ld_out, stream_ptr(locd.device)),
                  "vaemdl_dlogistic_bwd")
        return dloc, dls, None, None, None, None, None, None


def dlogistic_log_prob(loc, logscale, x, low=-1.0, high=1.0, levels=256.0) -> torch.Tensor:
    """Element-wise plain discretized-logistic log-prob (utils/discretized_logistic.py:35-78)."""
    n_event = min(3, loc.dim())
    return _DlFn.apply(loc, logscale, x, low, high, levels, n_event, "elem")


def dlogistic_log_likelihood(loc, logscale, x, low=-1.0, high=1.0, levels=256.0, n_event_dims: int = 3,
                             dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """``reduce_sum(log_prob(x), last n_event_dims axes)`` (models/loss.py:32) without the element-wise tensor."""
    return _DlFn.apply(loc, logscale, x, low, high, levels, n_event_dims, "image64" if dtype == torch.float64 else "image")


def dlogistic_iwae_forward(loc, logscale, x, extra: Optional[torch.Tensor] = None, low=-1.0, high=1.0, levels=256.0,
                           b_total: int = 0, n_event_dims: int = 3):
    """Plain discretized-logistic forward fused with the IWAE tail (``vaemdl_dlogistic_iwae_fwd``, two launches):
    ``loc``/``logscale`` ``[S, B, *event]``, ``x [B, *event]``, ``extra [S,B]`` = every other term of ``log_w``.
    Returns ``(lpxz float64 [S,B], log_w, lme_b [B], elbo [1], g_ll [S,B])`` (models/loss.py:32-37, models/model06.py:45-50)."""
    if loc.shape != logscale.shape:
        loc, logscale = torch.broadcast_tensors(loc, logscale)
    _abi.require_cuda(loc, "loc")
    if loc.dim() != 2 + n_event_dims:
        raise ValueError("loc must be [S, B, *event]")
    locd, lsd, C, ld = _dl_layout(loc, logscale)
    S, B = loc.shape[:2]
    ev = tuple(loc.shape[2:])
    D = int(math.prod(ev))
    xd, x_dtype, x_batch = _prep_x(x, ev, "x")
    if x_batch not in (1, B):
        raise ValueError(f"x must hold {B} images (or one), got {x_batch}")
    ex = dense_f32(extra, "extra").reshape(S, B) if extra is not None else None
    L = lib()
    dev = loc.device
    ll64 = torch.empty((S, B), device=dev, dtype=torch.float64)
    log_w = torch.empty((S, B), device=dev, dtype=torch.float32)
    lme_b = torch.empty(B, device=dev, dtype=torch.float32)
    elbo = torch.empty(1, device=dev, dtype=torch.float32)
    g_ll = torch.empty((S, B), device=dev, dtype=torch.float32)
    ws_bytes = L.vaemdl_dlogistic_workspace_bytes(S * B, D)
    ws = torch.empty((ws_bytes + 7) // 8, device=dev, dtype=torch.float64)
    with _abi.on_device(dev):
        check(L.vaemdl_dlogistic_iwae_fwd(ptr(locd), ptr(lsd), C, ld, ptr(xd), x_dtype, S, B, int(b_total), x_batch, D,
                                          float(low), float(high), float(levels), ptr(ex), None, ptr(ll64), ptr(log_w),
                                          ptr(lme_b), ptr(elbo), ptr(g_ll), ptr(ws), ws_bytes, stream_ptr(dev)),
              "vaemdl_dlogistic_iwae_fwd")
    return ll64, log_w, lme_b, elbo, g_ll


def dlogistic_iwae_step(loc, logscale, x, extra: Optional[torch.Tensor] = None, low=-1.0, high=1.0, levels=256.0,
                        b_total: int = 0, n_event_dims: int = 3):
    """Plain discretized-logistic forward + IWAE finish + gradient in ONE call (``vaemdl_dlogistic_iwae_step``): a single
    cooperative kernel launch for the image shapes of models 03/04/06 (the parameters are read once, the unscaled
    derivatives wait in shared memory for the importance weights), three launches otherwise.  Arguments as
    ``dlogistic_iwae_forward``.  Returns ``(lpxz float64 [S,B], log_w, lme_b, elbo, g_ll, dloc, dlogscale, launches)``;
    for the two halves of an un-split ``[..., 2C]`` tensor the two gradients are views of one ``[..., 2C]`` buffer."""
    if loc.shape != logscale.shape:
        loc, logscale = torch.broadcast_tensors(loc, logscale)
    _abi.require_cuda(loc, "loc")
    if loc.dim() != 2 + n_event_dims:
        raise ValueError("loc must be [S, B, *event]")
    locd, lsd, C, ld = _dl_layout(loc, logscale)
    shape = tuple(loc.shape)
    S, B = shape[:2]
    ev = shape[2:]
    D = int(math.prod(ev))
    xd, x_dtype, x_batch = _prep_x(x, ev, "x")
    if x_batch not in (1, B):
        raise ValueError(f"x must hold {B} images (or one), got {x_batch}")
    ex = dense_f32(extra, "extra").reshape(S, B) if extra is not None else None
    L = lib()
    dev = loc.device
    ll64 = torch.empty((S, B), device=dev, dtype=torch.float64)
    log_w = torch.empty((S, B), device=dev, dtype=torch.float32)
    lme_b = torch.empty(B, device=dev, dtype=torch.float32)
    elbo = torch.empty(1, device=dev, dtype=torch.float32)
    g_ll = torch.empty((S, B), device=dev, dtype=torch.float32)
    if ld == 2 * C:
        both = torch.empty(shape[:-1] + (2 * C,), device=dev, dtype=torch.float32)
        dloc, dls, ld_out = both[..., :C], both[..., C:], 2 * C
        p_loc, p_ls = both.data_ptr(), both.data_ptr() + 4 * C
    else:
        dloc = torch.empty(shape, device=dev, dtype=torch.float32)
        dls = torch.empty(shape, device=dev, dtype=torch.float32)
        ld_out, p_loc, p_ls = C, dloc.data_ptr(), dls.data_ptr()
    ws_bytes = L.vaemdl_dlogistic_workspace_bytes(S * B, D)
    ws = torch.empty((ws_bytes + 7) // 8, device=dev, dtype=torch.float64)
    n_launch = ctypes.c_int(0)
    with _abi.on_device(dev):
        check(L.vaemdl_dlogistic_iwae_step(ptr(locd), ptr(lsd), C, ld, ptr(xd), x_dtype, S, B, int(b_total), x_batch, D,
                                           float(low), float(high), float(levels), ptr(ex), None, ptr(ll64), ptr(log_w),
                                           ptr(lme_b), ptr(elbo), ptr(g_ll), ctypes.c_void_p(p_loc), ctypes.c_void_p(p_ls),
                                           ld_out, ptr(ws), ws_bytes, stream_ptr(dev), ctypes.byref(n_launch)),
              "vaemdl_dlogistic_iwae_step")
    return ll64, log_w, lme_b, elbo, g_ll, dloc, dls, n_launch.value


def dlogistic_backward(loc, logscale, x, g_image: torch.Tensor, low=-1.0, high=1.0, levels=256.0, n_event_dims: int = 3):
    """The plain-DL gradient kernel on its own: ``d/d(loc, logscale)`` of ``sum(g_image * ll_image)``.  For the two halves
    of an un-split ``[..., 2C]`` tensor the two gradients are views of one ``[..., 2C]`` buffer."""
    if loc.shape != logscale.shape:
        loc, logscale = torch.broadcast_tensors(loc, logscale)
    _abi.require_cuda(loc, "loc")
    locd, lsd, C, ld = _dl_layout(loc, logscale)
    shape = tuple(loc.shape)
    ev = shape[len(shape) - n_event_dims:]
    lead = shape[:len(shape) - n_event_dims]
    D = int(math.prod(ev))
    n_img = int(math.prod(lead)) if lead else 1
    xd, x_dtype, x_batch = _prep_x(x, ev, "x")
    _check_batch(n_img, x_batch, "backward")
    gi = dense_f32(g_image, "g_image")
    if ld == 2 * C:
        both = torch.empty(shape[:-1] + (2 * C,), device=locd.device, dtype=torch.float32)
        dloc, dls, ld_out = both[..., :C], both[..., C:], 2 * C
        p_loc, p_ls = both.data_ptr(), both.data_ptr() + 4 * C
    else:
        dloc = torch.empty(shape, device=locd.device, dtype=torch.float32)
        dls = torch.empty(shape, device=locd.device, dtype=torch.float32)
        ld_out, p_loc, p_ls = C, dloc.data_ptr(), dls.data_ptr()
    import ctypes
    with _abi.on_device(locd.device):
        check(lib().vaemdl_dlogistic_bwd(ptr(locd), ptr(lsd), C, ld, ptr(xd), x_dtype, n_img, x_batch, D, float(low),
                                         float(high), float(levels), ptr(gi), None, ctypes.c_void_p(p_loc),
                                         ctypes.c_void_p(p_ls), ld_out, stream_ptr(locd.device)), "vaemdl_dlogistic_bwd")
    return dloc, dls


# --------------------------------------------------------------------------------------------------
# log-mean-exp / IWAE tail
# --------------------------------------------------------------------------------------------------
class _LmeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, log_w2d):
        S, B = log_w2d.shape
        out = torch.empty(B, device=log_w2d.device, dtype=torch.float32)
        L = lib()
        fn = L.vaemdl_logmeanexp_fwd_f64 if log_w2d.dtype == torch.float64 else L.vaemdl_logmeanexp_fwd
        with _abi.on_device(log_w2d.device):
            check(fn(ptr(log_w2d), S, B, ptr(out), stream_ptr(log_w2d.device)), "vaemdl_logmeanexp_fwd")
        ctx.save_for_backward(log_w2d)
        return out

    @staticmethod
    def backward(ctx, g):
        (log_w2d,) = ctx.saved_tensors
        S, B = log_w2d.shape
        g = dense_f32(g, "grad")
        d = torch.empty(log_w2d.shape, device=log_w2d.device, dtype=torch.float32)
        L = lib()
        fn = L.vaemdl_logmeanexp_bwd_f64 if log_w2d.dtype == torch.float64 else L.vaemdl_logmeanexp_bwd
        with _abi.on_device(log_w2d.device):
            check(fn(ptr(log_w2d), ptr(g), S, B, ptr(d), stream_ptr(log_w2d.device)), "vaemdl_logmeanexp_bwd")
        return d.to(log_w2d.dtype)


def _dense_f32_or_f64(t: torch.Tensor, name: str) -> torch.Tensor:
    if t.dtype == torch.float64:
        _abi.require_cuda(t, name)
        return t.contiguous()
    return dense_f32(t, name)


def logmeanexp(log_w: torch.Tensor, axis: int) -> torch.Tensor:
    """``log(mean(exp(log_w), axis))`` computed stably (utils/utils.py:9-11), differentiable.  float64 input is
    reduced with float64 differences; the result is float32 either way."""
    _abi.require_cuda(log_w, "log_w")
    axis = axis % log_w.dim()
    moved = log_w.movedim(axis, 0)
    rest = tuple(moved.shape[1:])
    flat = _dense_f32_or_f64(moved.reshape(moved.shape[0], -1), "log_w")
    return _LmeFn.apply(flat).reshape(rest)


def iwae_tail(ll: torch.Tensor, extra: Optional[torch.Tensor] = None, b_total: int = 0):
    """Fused IWAE tail on ``ll [S,B]`` (float32 or float64) + ``extra [S,B]``: returns ``(log_w, lme_b, elbo, g_ll)``
    (all float32) where ``g_ll = d(-elbo)/d ll`` (models/loss.py:34-37).  ``b_total``: size of the whole batch when
    ``ll`` holds one rank's shard of it (0 = this is the whole batch).  Not recorded by autograd."""
    ll = _dense_f32_or_f64(ll, "ll")
    S, B = ll.shape
    ex = dense_f32(extra, "extra") if extra is not None else None
    log_w = torch.empty((S, B), device=ll.device, dtype=torch.float32)
    lme_b = torch.empty(B, device=ll.device, dtype=torch.float32)
    elbo = torch.empty(1, device=ll.device, dtype=torch.float32)
    g_ll = torch.empty((S, B), device=ll.device, dtype=torch.float32)
    is64 = ll.dtype == torch.float64
    with _abi.on_device(ll.device):
        check(lib().vaemdl_iwae_tail(None if is64 else ptr(ll), ptr(ll) if is64 else None, ptr(ex), S, B, int(b_total), ptr(log_w),
                                     ptr(lme_b), ptr(elbo), ptr(g_ll), stream_ptr(ll.device)), "vaemdl_iwae_tail")
    return log_w, lme_b, elbo, g_ll


# --------------------------------------------------------------------------------------------------
# samplers
# --------------------------------------------------------------------------------------------------
def modl_sample(params: torch.Tensor, u_mix: torch.Tensor, u_log: torch.Tensor, variant: int = _abi.SAMPLE_OPENAI,
                out_range: int = _abi.RANGE_SYM, want_quantised: bool = False, want_index: bool = False,
                clip=(-1.0, 1.0)):
    """Explicit-noise MoDL sampler.  ``params [..., H, W, 10M]``; the noise may carry extra leading dims ``[n..., ...]``
    (``u_mix [n..., ..., H, W, M]``): the parameters are then re-used for every ``n`` without being tiled.
    Returns ``x [n..., ..., H, W, 3]`` float32 (+ uint8 quantised values, + uint8 mixture indices)."""
    p = dense_f32(params, "parameters")
    H, W, C10 = p.shape[-3:]
    M = C10 // 10
    lead = tuple(p.shape[:-3])
    n_img = int(math.prod(lead)) if lead else 1
    um = dense_f32(u_mix, "u_mix")
    ul = dense_f32(u_log, "u_log") if u_log is not None else None
    if ul is None and variant != _abi.SAMPLE_PLAIN:
        raise ValueError("u_log may only be omitted for the plain variant (mean of the selected locations)")
    per_rep = n_img * H * W * M
    if um.shape[-1] != M or um.numel() % per_rep:
        raise ValueError("u_mix must have shape [n..., ..., H, W, n_mix]")
    n_rep = um.numel() // per_rep
    out_lead = tuple(um.shape[:-3])
    if int(math.prod(out_lead)) != n_rep * n_img:
        raise ValueError("u_mix leading dims must be [n...] + parameters.shape[:-3]")
    want = n_rep * n_img * H * W * 3 * (1 if variant == _abi.SAMPLE_OPENAI else M)
    if ul is not None and ul.numel() != want:
        raise ValueError("u_log must have shape [n..., ..., H, W, 3] (openai) or [n..., ..., H, W, 3, n_mix] (mdl)")
    x = torch.empty(out_lead + (H, W, 3), device=p.device, dtype=torch.float32)
    xq = torch.empty(out_lead + (H, W, 3), device=p.device, dtype=torch.uint8) if want_quantised else None
    idx = torch.empty(out_lead + (H, W), device=p.device, dtype=torch.uint8) if want_index else None
    with _abi.on_device(p.device):
        if variant == _abi.SAMPLE_PLAIN and tuple(float(v) for v in clip) != (-1.0, 1.0):
            # the un-conditioned class built with its own low / high (utils/mdl_plain.py:18, discretized_logistic.py:83)
            check(lib().vaemdl_modl_plain_sample(ptr(p), ptr(um), ptr(ul), float(clip[0]), float(clip[1]), out_range, n_rep,
                                                 n_img, H, W, M, ptr(x), ptr(xq), ptr(idx), stream_ptr(p.device)),
                  "vaemdl_modl_plain_sample")
        else:
            check(lib().vaemdl_modl_sample(ptr(p), ptr(um), ptr(ul), variant, out_range, n_rep, n_img, H, W, M, ptr(x),
                                           ptr(xq), ptr(idx), stream_ptr(p.device)), "vaemdl_modl_sample")
    outs = [x]
    if want_quantised:
        outs.append(xq)
    if want_index:
        outs.append(idx)
    return outs[0] if len(outs) == 1 else tuple(outs)


def dlogistic_sample(loc: torch.Tensor, logscale: torch.Tensor, u: torch.Tensor, low=-1.0, high=1.0) -> torch.Tensor:
    """``clip(loc + exp(logscale) * (log u - log(1-u)), low, high)`` (utils/discretized_logistic.py:80-85)."""
    if loc.shape != logscale.shape:
        loc, logscale = torch.broadcast_tensors(loc, logscale)
    _abi.require_cuda(loc, "loc")
    locd, lsd, C, ld = _dl_layout(loc, logscale)
    n_param = loc.numel()
    ud = dense_f32(u, "u")
    if ud.numel() % n_param:
        raise ValueError("u must have shape [n..., *loc.shape]")
    out = torch.empty(ud.shape, device=loc.device, dtype=torch.float32)
    reps = ud.numel() // n_param
    with _abi.on_device(loc.device):
        for r in range(reps):  # leading sample dims re-use the same parameters
            import ctypes
            off = r * n_param * 4
            check(lib().vaemdl_dlogistic_sample(ptr(locd), ptr(lsd), C, ld, ctypes.c_void_p(ud.data_ptr() + off), n_param,
                                                float(low), float(high), ctypes.c_void_p(out.data_ptr() + off),
                                                stream_ptr(loc.device)), "vaemdl_dlogistic_sample")
    return out


# --------------------------------------------------------------------------------------------------
# latent-side terms of log_w, and the whole IWAE objective after the networks in 3 launches (+ 2 backward)
# --------------------------------------------------------------------------------------------------
def _pack_terms(terms):
    """terms: list of ``(z [S,B,D], loc | None, scale | None, weight)``; loc / scale ``[B,D]`` or ``[S,B,D]``.
    Returns (ctypes array, tensors kept alive, S, B)."""
    n = len(terms)
    if not 1 <= n <= _abi.MAX_LATENT_TERMS:
        raise ValueError(f"between 1 and {_abi.MAX_LATENT_TERMS} terms")
    arr = (_abi.LatentTerm * n)()
    keep = []
    S, B = terms[0][0].shape[0], terms[0][0].shape[1]
    for i, (z, loc, scale, w) in enumerate(terms):
        zd = dense_f32(z, "z")
        if zd.dim() != 3 or zd.shape[0] != S or zd.shape[1] != B:
            raise ValueError("every z must be [S, B, D]")
        D = zd.shape[2]
        ld = sd = None
        pps = 0
        if loc is not None:
            ld, sd = dense_f32(loc, "loc"), dense_f32(scale, "scale")
            if tuple(ld.shape) == (S, B, D) and tuple(sd.shape) == (S, B, D):
                pps = 1
            elif tuple(ld.shape) != (B, D) or tuple(sd.shape) != (B, D):
                raise ValueError("loc / scale must be [B, D] or [S, B, D]")
        arr[i].z = zd.data_ptr()
        arr[i].loc = ld.data_ptr() if ld is not None else None
        arr[i].scale = sd.data_ptr() if sd is not None else None
        arr[i].D = D
        arr[i].params_per_sample = pps
        arr[i].weight = float(w)
        keep.append((zd, ld, sd))
    return arr, keep, S, B


def latent_terms(terms, extra_in: Optional[torch.Tensor] = None):
    """``extra [S,B] = extra_in + sum_t w_t * sum_d log N(z_t; loc_t, scale_t)`` and the per-term sums ``[n,S,B]``
    (models/loss.py:28-34, models/model06.py:40-47) in ONE launch.  Not recorded by autograd (see ``fused_iwae_loss``)."""
    arr, keep, S, B = _pack_terms(terms)
    dev = keep[0][0].device
    extra = torch.empty((S, B), device=dev, dtype=torch.float32)
    sums = torch.empty((len(terms), S, B), device=dev, dtype=torch.float32)
    ex = dense_f32(extra_in, "extra_in") if extra_in is not None else None
    with _abi.on_device(dev):
        check(lib().vaemdl_latent_terms_fwd(arr, len(terms), S, B, ptr(ex), ptr(extra), ptr(sums), stream_ptr(dev)),
              "vaemdl_latent_terms_fwd")
    return extra, sums


def latent_terms_backward(terms, g_extra: torch.Tensor, share_dz=()):
    """Gradients of ``sum(g_extra * extra)`` w.r.t. every term's z, loc, scale (ONE launch).  ``share_dz``: pairs
    ``(i, j)`` of terms that are densities of the SAME z tensor -- term j then accumulates into term i's dz buffer.
    Returns lists ``dz, dloc, dscale`` (``None`` for a standard-normal term's parameters, and for dz[j] of a shared pair)."""
    import ctypes
    arr, keep, S, B = _pack_terms(terms)
    n = len(terms)
    dev = keep[0][0].device
    g = dense_f32(g_extra, "g_extra")
    dz = [torch.empty_like(k[0]) for k in keep]
    for i, j in share_dz:
        dz[j] = dz[i]
    dloc = [torch.empty_like(k[1]) if k[1] is not None else None for k in keep]
    dsc = [torch.empty_like(k[2]) if k[2] is not None else None for k in keep]
    PP = ctypes.c_void_p * n
    mk = lambda lst: PP(*[t.data_ptr() if t is not None else None for t in lst])  # noqa: E731
    with _abi.on_device(dev):
        check(lib().vaemdl_latent_terms_bwd(arr, n, S, B, ptr(g), mk(dz), mk(dloc), mk(dsc), stream_ptr(dev)),
              "vaemdl_latent_terms_bwd")
    shared = {j for _, j in share_dz}
    return [None if i in shared else dz[i] for i in range(n)], dloc, dsc


def _expand_param(t: torch.Tensor, S: int, B: int, D: int):
    """A Normal parameter broadcast to [B, D] when it does not depend on the sample axis, else to [S, B, D]."""
    t = t.float()
    if t.dim() < 3 or t.shape[0] == 1:
        tb = t.reshape(t.shape[-2:]) if t.dim() >= 3 else t
        return tb.expand(B, D).contiguous()
    return t.expand(S, B, D).contiguous()


def _reduce_to(g: torch.Tensor, shape) -> torch.Tensor:
    """Sums the gradient of an expanded Normal parameter back to the parameter's own shape (which may carry leading
    singleton dims the expansion dropped, e.g. ``[1,B,D]`` -> ``[B,D]``)."""
    shape = tuple(shape)
    if g.dim() < len(shape):
        g = g.reshape((1,) * (len(shape) - g.dim()) + tuple(g.shape))
    return g.sum_to_size(shape) if len(shape) else g.sum()


class _FusedIwaeFn(torch.autograd.Function):
    """The IWAE objective after the networks (models/loss.py:26-46, models/model06.py:38-55): latent terms ->
    observation-model forward -> fused finish; backward: observation-model gradient kernel + latent-term gradient kernel.

    ``term_meta``: one ``(z_index, weight, param_index)`` per Normal term (``param_index = -1``: standard normal);
    ``tensors`` = the distinct z tensors, then ``loc, scale`` of every parametrised term in order."""

    @staticmethod
    def forward(ctx, kind, meta, x, term_meta, n_z, p0, p1, *tensors):
        zs = [dense_f32(t, "z") for t in tensors[:n_z]]
        S, B = zs[0].shape[0], zs[0].shape[1]
        raw = tensors[n_z:]
        terms, expanded = [], []
        for zi, w, pi in term_meta:
            D = zs[zi].shape[2]
            if pi < 0:
                terms.append((zs[zi], None, None, w))
            else:
                le, se = _expand_param(raw[2 * pi], S, B, D), _expand_param(raw[2 * pi + 1], S, B, D)
                expanded += [le, se]
                terms.append((zs[zi], le, se, w))
        extra, sums = latent_terms(terms)
        stats = None
        if kind == "modl":
            # n_mix 5 / 30: the forward kernel leaves the per-pixel mixture sums for a one-pass gradient kernel (the
            # library ignores them for any other n_mix, so they are not even allocated then)
            stat_widths = (60, 80, 100, 120, 140, 160, 180, 200, 240, 280, 300, 320, 360, 400, 480, 500, 560, 600, 640) if p0.dtype == torch.bfloat16 else (50, 300)
            want = bool(ctx.needs_input_grad[5]) and p0.shape[-1] in stat_widths and not meta["plain"]
            out = modl_iwae_forward(p0, x, extra, 0, meta["x_range"], meta["edge_mode"], meta["plain"], want_stats=want)
            ll64, log_w, lme_b, elbo, g_ll = out[:5]
            stats = out[5] if want else None
        else:
            ll64, log_w, lme_b, elbo, g_ll = dlogistic_iwae_forward(p0, p1, x, extra, meta["low"], meta["high"],
                                                                    meta["levels"])
        ctx.kind, ctx.meta_, ctx.term_meta, ctx.n_z = kind, meta, term_meta, n_z
        ctx.raw_shapes = [t.shape for t in raw]
        ctx.has_p1 = p1 is not None
        ctx.pix_stats = stats   # (not an input or output of the function: plain attribute)
        ctx.save_for_backward(x, p0, p1 if p1 is not None else p0, g_ll, *zs, *expanded)
        lpxz = ll64.float()
        loss = -elbo.reshape(())
        ctx.mark_non_differentiable(lpxz, sums, lme_b)
        return loss, lpxz, sums, lme_b

    @staticmethod
    def backward(ctx, g_loss, *_unused):
        x, p0, p1, g_ll = ctx.saved_tensors[:4]
        zs = ctx.saved_tensors[4:4 + ctx.n_z]
        expanded = ctx.saved_tensors[4 + ctx.n_z:]
        g = g_ll * g_loss                                                     # d loss / d lpxz = d loss / d extra
        meta = ctx.meta_
        dp0 = dp1 = None
        if ctx.kind == "modl":
            if ctx.needs_input_grad[5]:
                dp0 = modl_backward(p0, x, g_image=g, x_range=meta["x_range"], edge_mode=meta["edge_mode"],
                                    plain=meta["plain"], pix_stats=None if meta["plain"] else ctx.pix_stats)
        elif ctx.needs_input_grad[5] or ctx.needs_input_grad[6]:
            dp0, dp1 = dlogistic_backward(p0, p1, x, g, meta["low"], meta["high"], meta["levels"])
        n_t = len(ctx.needs_input_grad) - 7
        d_tensors = [None] * n_t
        if any(ctx.needs_input_grad[7:]):
            terms, first_of_z, share = [], {}, []
            for t, (zi, w, pi) in enumerate(ctx.term_meta):
                if pi < 0:
                    terms.append((zs[zi], None, None, w))
                else:
                    terms.append((zs[zi], expanded[2 * pi], expanded[2 * pi + 1], w))
                if zi in first_of_z:
                    share.append((first_of_z[zi], t))
                else:
                    first_of_z[zi] = t
            dzs, dlocs, dscs = latent_terms_backward(terms, g, share_dz=tuple(share))
            for zi, t in first_of_z.items():
                d_tensors[zi] = dzs[t]
            for t, (zi, w, pi) in enumerate(ctx.term_meta):
                if pi >= 0:
                    d_tensors[ctx.n_z + 2 * pi] = _reduce_to(dlocs[t], ctx.raw_shapes[2 * pi])
                    d_tensors[ctx.n_z + 2 * pi + 1] = _reduce_to(dscs[t], ctx.raw_shapes[2 * pi + 1])
        if not ctx.has_p1:
            dp1 = None
        return (None, None, None, None, None, dp0, dp1, *d_tensors)


def fused_iwae_loss(kind: str, meta: dict, x, p0, p1, normal_terms):
    """``normal_terms``: list of ``(z [S,B,D], loc | None, scale | None, weight)`` -- the Normal densities that make up
    ``log_w - lpxz`` (a term with ``loc is None`` is the standard normal); terms may share a z tensor.
    Returns ``(loss, lpxz [S,B], term_sums [n_terms,S,B], lme_b [B])``; ``loss`` is differentiable w.r.t. the
    observation-model parameters, every z and every loc / scale.  3 launches forward, 2 backward."""
    z_list, term_meta, params = [], [], []
    for z, loc, scale, w in normal_terms:
        zi = next((i for i, t in enumerate(z_list) if t is z), None)
        if zi is None:
            z_list.append(z)
            zi = len(z_list) - 1
        if loc is None:
            term_meta.append((zi, float(w), -1))
        else:
            term_meta.append((zi, float(w), len(params) // 2))
            params += [loc, scale]
    return _FusedIwaeFn.apply(kind, meta, x, tuple(term_meta), len(z_list), p0, p1, *z_list, *params)
