"""Shared helpers for the GPU parity tests (test infrastructure; imports the oracle as the checker)."""
import os

import numpy as np
import torch

import oracle as O

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
NAMES = ["logit", "muR", "sR", "kR", "muG", "sG", "kG", "muB", "sB", "kB"]

# north_star tolerances
LL_RTOL = 1e-5     # per-image log-likelihood / bits-per-dim, relative
GRAD_RTOL = 1e-4   # gradients, relative (normwise, per parameter group and overall)


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def relnorm(a, b):
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-300)).item()


def assert_ll_close(ll_gpu, ll_ref, rtol=LL_RTOL):
    a = ll_gpu.detach().double().cpu()
    b = torch.as_tensor(ll_ref).double()
    err = ((a - b).abs() / b.abs().clamp_min(1e-30)).max().item()
    assert err <= rtol, f"per-image log-likelihood: max relative error {err:.3e} > {rtol:g}"


def assert_grad_close(g_gpu, g_ref, M, rtol=GRAD_RTOL):
    """normwise per parameter group and overall, plus element-wise |d| <= rtol*|ref| + rtol*max|ref|."""
    a = g_gpu.detach().double().cpu()
    b = torch.as_tensor(g_ref).double()
    assert a.shape == b.shape
    overall = relnorm(a, b)
    assert overall <= rtol, f"gradient (overall): normwise relative error {overall:.3e} > {rtol:g}"
    for j, nme in enumerate(NAMES):
        aj, bj = a[..., j * M:(j + 1) * M], b[..., j * M:(j + 1) * M]
        if bj.norm().item() <= 1e-12 * b.norm().item():  # identically-zero group (e.g. the logit of a 1-component mixture)
            assert aj.norm().item() <= rtol * b.norm().item(), f"gradient group {nme}: expected ~0"
            continue
        e = relnorm(aj, bj)
        assert e <= rtol, f"gradient group {nme}: normwise relative error {e:.3e} > {rtol:g}"
    bound = rtol * b.abs() + rtol * b.abs().max()
    worst = ((a - b).abs() - bound).max().item()
    assert worst <= 0, f"gradient element-wise bound exceeded by {worst:.3e}"


def canonical(seed, S, B, H, W, M):
    g = torch.Generator().manual_seed(seed)
    x_u8 = torch.randint(0, 256, (B, H, W, 3), dtype=torch.uint8, generator=g)
    params = torch.randn(S, B, H, W, 10 * M, generator=g)
    return params, x_u8, g


def trained_like(seed, S, B, H, W, M):
    g = torch.Generator().manual_seed(seed)
    x_u8 = torch.randint(0, 256, (B, H, W, 3), dtype=torch.uint8, generator=g)
    p = torch.randn(S, B, H, W, 10 * M, generator=g)
    rest = p[..., M:].reshape(S, B, H, W, 3, 3 * M).clone()
    rest[..., :M] = torch.rand(S, B, H, W, 3, M, generator=g) * 2 - 1
    rest[..., M:2 * M] = torch.randn(S, B, H, W, 3, M, generator=g) - 3.0
    return torch.cat([p[..., :M], rest.reshape(S, B, H, W, 9 * M)], -1).contiguous(), x_u8, g


def threshold_ambiguous(params64, x64, rel=1e-3):
    """Images holding a sub-pixel whose CDF difference sits within `rel` of the 1e-5 branch threshold: there the
    reference's own branch choice depends on float32 rounding, so exact parity is not defined."""
    x = x64 * 2 - 1
    loc, logscale, _ = O.ref._mdl_autoregressive_params(params64, x)
    inv = torch.exp(-logscale)
    xx = x[..., None]
    prob = torch.sigmoid((xx - loc + 1 / 255) * inv) - torch.sigmoid((xx - loc - 1 / 255) * inv)
    edge = ((xx <= -1) | (xx >= 1)).expand_as(prob)
    amb = ((prob - 1e-5).abs() < rel * 1e-5) & ~edge
    return amb.flatten(-4).any(-1)  # [..., B]


def oracle_ll_and_grad(params, x_u8, g_image):
    p64 = params.double().requires_grad_(True)
    x64 = O.normalize_u8(x_u8, torch.float64)
    lp = O.modl_log_prob(p64, x64)
    ll = lp.sum((-1, -2, -3))
    (ll * g_image.double()).sum().backward()
    return lp.detach()[..., 0], ll.detach(), p64.grad
