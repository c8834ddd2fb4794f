"""Seeded random shapes through every MoDL kernel family (tile, pixel-pair, run-time tile, one-launch step, bf16) and the
plain discretized logistic, against the float64 oracle: ragged tiles, images smaller than a tile, odd n_mix, both edge
modes, uint8 and float observations."""
import random

import pytest
import torch

import oracle as O
from util import GRAD_RTOL, LL_RTOL, relnorm, threshold_ambiguous, trained_like

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _cases(n, seed):
    rng = random.Random(seed)
    out = []
    for _ in range(n):
        M = rng.choice([1, 2, 3, 4, 5, 5, 6, 7, 8, 9, 10, 10, 11, 12, 13, 14, 16, 17, 20, 24, 25, 30, 33, 40, 48, 64])
        H, W = rng.randint(1, 12), rng.randint(1, 12)
        S, B = rng.randint(1, 4), rng.randint(1, 4)
        out.append((S, B, H, W, M, rng.random() < 0.5, rng.random() < 0.5, rng.randint(0, 10 ** 6)))
    return out


@pytest.mark.parametrize("S,B,H,W,M,openai,u8,seed", _cases(48, 20261018))
def test_random_shapes_modl(built_lib, S, B, H, W, M, openai, u8, seed):
    from vae_mdl_b200 import _abi, functional as F
    params, x_u8, g = trained_like(seed, S, B, H, W, M)
    x_u8.view(-1)[::7] = 0
    x_u8.view(-1)[3::11] = 255
    g_image = torch.randn(S, B, generator=g)
    x01_32 = O.normalize_u8(x_u8, torch.float32)
    x64 = x01_32.double()
    p64 = params.double().requires_grad_(True)
    if openai:
        lp = O.modl_openai_iwae_log_prob(p64, x64)[..., 0]
    else:
        lp = O.modl_log_prob(p64, x64)[..., 0]
    ll64 = lp.sum((-1, -2))
    (ll64 * g_image.double()).sum().backward()
    ok = ~threshold_ambiguous(params.double(), x64)
    if not bool(ok.any()):
        pytest.skip("every image of this draw holds a threshold-ambiguous sub-pixel")
    edge = _abi.EDGE_OPENAI if openai else _abi.EDGE_MDL
    pd = params.to(DEV)
    xd = x_u8.to(DEV) if u8 else x01_32.to(DEV)
    ll = F.modl_log_likelihood(pd, xd, _abi.RANGE_UNIT, edge, torch.float64).cpu()
    assert ((ll - ll64.detach()).abs() / ll64.detach().abs().clamp_min(1e-3))[ok].max().item() <= LL_RTOL
    lpd = F.modl_log_prob(pd, xd, _abi.RANGE_UNIT, edge).cpu().double()
    assert (((lpd - lp.detach()).abs() - 1e-6 * lp.detach().abs()).flatten(2).amax(-1))[ok].max().item() < 5e-5
    dp = F.modl_backward(pd, xd, g_image=g_image.to(DEV), x_range=_abi.RANGE_UNIT, edge_mode=edge).cpu().double()
    assert relnorm(dp[ok], p64.grad[ok]) <= GRAD_RTOL
    # one-call step (one cooperative launch when eligible) and the bf16 interface on the same draw
    extra = (ll64.detach().mean(0, keepdim=True) - ll64.detach()).float()
    out = F.modl_iwae_step(pd, xd, extra.to(DEV), edge_mode=edge)
    assert ((out[0].cpu() - ll64.detach()).abs() / ll64.detach().abs().clamp_min(1e-3))[ok].max().item() <= LL_RTOL
    pb = params.bfloat16()
    llb = F.modl_log_likelihood(pb.to(DEV), xd, _abi.RANGE_UNIT, edge, torch.float64)
    llw = F.modl_log_likelihood(pb.float().to(DEV), xd, _abi.RANGE_UNIT, edge, torch.float64)
    assert ((llb - llw).abs() / llw.abs().clamp_min(1e-3)).max().item() <= 1e-6


@pytest.mark.parametrize("S,B,H,W,M,interleaved,u8,seed", _cases(16, 77))
def test_random_shapes_plain_dl(built_lib, S, B, H, W, M, interleaved, u8, seed):
    from vae_mdl_b200 import functional as F
    g = torch.Generator().manual_seed(seed)
    both = torch.randn(S, B, H, W, 6, generator=g)
    both[..., :3] = torch.rand(S, B, H, W, 3, generator=g)
    both[..., 3:] -= 2.0
    x_u8 = torch.randint(0, 256, (B, H, W, 3), dtype=torch.uint8, generator=g)
    x_u8.view(-1)[::5] = 0
    x_u8.view(-1)[2::9] = 255
    loc64 = both[..., :3].double().requires_grad_(True)
    ls64 = both[..., 3:].double().requires_grad_(True)
    x64 = O.normalize_u8(x_u8, torch.float32).double()
    ll64 = O.dlogistic_log_prob(x64, loc64, ls64, 0.0, 1.0, 256.0).sum((-1, -2, -3))
    extra = (ll64.detach().mean(0, keepdim=True) - ll64.detach()).float()
    loss64 = -O.logmeanexp(ll64 + extra.double(), 0).mean()
    loss64.backward()
    bd = both.to(DEV)
    loc, ls = (bd[..., :3], bd[..., 3:]) if interleaved else (bd[..., :3].contiguous(), bd[..., 3:].contiguous())
    xd = x_u8.to(DEV) if u8 else O.normalize_u8(x_u8, torch.float32).to(DEV)
    out = F.dlogistic_iwae_step(loc, ls, xd, extra.to(DEV), 0.0, 1.0, 256.0)
    assert ((out[0].cpu() - ll64.detach()).abs() / ll64.detach().abs().clamp_min(1e-3)).max().item() <= LL_RTOL
    assert abs(-out[3].item() - loss64.item()) <= LL_RTOL * abs(loss64.item())
    assert relnorm(out[5], loc64.grad) <= GRAD_RTOL and relnorm(out[6], ls64.grad) <= GRAD_RTOL


@pytest.mark.parametrize("M", [3, 5, 10, 12, 16, 25])
def test_broadcast_single_image_against_many_samples(built_lib, M):
    """models/model05.py:173: x WITHOUT a batch dim scored against [S, 1, H, W, 10M] parameters (x_batch = 1), every
    kernel family; the IWAE-class wrapper and the float-x path give the same numbers."""
    from vae_mdl_b200 import functional as F
    import vae_mdl_b200 as V
    S, H, W = 37, 9, 7
    params, x_u8, g = trained_like(4200 + M, S, 1, H, W, M)
    x1 = x_u8[0]                                             # [H, W, 3]
    p64 = params.double()
    x64 = O.normalize_u8(x1, torch.float32).double()
    lp = O.modl_log_prob(p64, x64)[..., 0]                   # broadcasting x over S and the batch dim of 1
    ll64 = lp.sum((-1, -2))
    ok = ~threshold_ambiguous(p64, x64[None])
    if not bool(ok.any()):
        pytest.skip("threshold-ambiguous draw")
    pd = params.to(DEV)
    ll_a = F.modl_log_likelihood(pd, x1.to(DEV), dtype=torch.float64).cpu()
    ll_b = F.modl_log_likelihood(pd, O.normalize_u8(x1, torch.float32).to(DEV), dtype=torch.float64).cpu()
    assert torch.equal(ll_a, ll_b)
    assert ((ll_a - ll64).abs() / ll64.abs().clamp_min(1e-3))[ok].max().item() <= LL_RTOL
    out = V.MixtureDiscretizedLogistic(pd).log_prob(O.normalize_u8(x1, torch.float32).to(DEV))
    assert list(out.shape) == [S, 1, H, W, 1]
    assert (((out[..., 0].cpu().double() - lp).abs() - 1e-6 * lp.abs()).flatten(2).amax(-1))[ok].max().item() < 5e-5
    g_image = torch.randn(S, 1, generator=g)
    q = p64.clone().requires_grad_(True)
    (O.modl_log_prob(q, x64)[..., 0].sum((-1, -2)) * g_image.double()).sum().backward()
    dp = F.modl_backward(pd, x1.to(DEV), g_image=g_image.to(DEV)).cpu().double()
    assert relnorm(dp[ok], q.grad[ok]) <= GRAD_RTOL
