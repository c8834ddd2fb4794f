"""GPU parity tests of ``PixelMixtureDiscretizedLogistic`` (utils/mdl_plain.py: means chained on the means)."""
import pytest
import torch

import oracle as O
from util import GRAD_RTOL, LL_RTOL, assert_grad_close, canonical, relnorm, trained_like

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def V(built_lib):
    import vae_mdl_b200
    return vae_mdl_b200


def threshold_ambiguous_plain(params64, x64, low, high, levels, rel=1e-3):
    """Pixels holding a sub-pixel whose CDF difference sits within `rel` of the 1e-5 branch threshold (the reference's own
    branch choice there depends on float32 rounding, cf. util.threshold_ambiguous).  -> bool [S,B,H,W]."""
    loc, logscale, _ = O.mdl_plain_get_mixture_params(params64)
    x = (x64 * 2 - 1)[..., None]
    dx = (high - low) / (levels - 1.0) / 2.0
    inv = torch.exp(-logscale)
    prob = torch.sigmoid((x - loc + dx) * inv) - torch.sigmoid((x - loc - dx) * inv)
    edge = ((x <= low) | (x >= high)).expand_as(prob)
    amb = ((prob - 1e-5).abs() < rel * 1e-5) & ~edge
    return amb.flatten(-2).any(-1)


def _oracle(params, x_u8, g_image):
    p64 = params.double().requires_grad_(True)
    x64 = O.normalize_u8(x_u8, torch.float64)
    lp = O.mdl_plain_log_prob(p64, x64)
    ll = lp.sum((-1, -2))
    (ll * g_image.double()).sum().backward()
    return lp.detach(), ll.detach(), p64.grad


@pytest.mark.parametrize("S,B,H,W,M", [(2, 3, 8, 8, 10), (2, 2, 8, 8, 5), (1, 2, 8, 8, 20), (1, 2, 8, 9, 30), (2, 1, 5, 7, 3),
                                        (1, 2, 4, 4, 13), (1, 1, 16, 16, 10), (2, 2, 3, 3, 7)])
@pytest.mark.parametrize("dist", ["canonical", "trained"])
def test_log_prob_and_gradient_vs_oracle(V, S, B, H, W, M, dist):
    make = canonical if dist == "canonical" else trained_like
    params, x_u8, g = make(900 + 13 * M + W, S, B, H, W, M)
    x_u8[0, 0, 0] = torch.tensor([0, 255, 128], dtype=torch.uint8)
    g_image = torch.randn(S, B, generator=g)
    lp64, ll64, grad64 = _oracle(params, x_u8, g_image)
    pd = params.to(DEV).requires_grad_(True)
    d = V.PixelMixtureDiscretizedLogistic(pd)
    lp = d.log_prob(O.normalize_u8(x_u8).to(DEV))
    assert lp.shape == (S, B, H, W)
    if dist == "canonical":
        assert (lp.detach().cpu().double() - lp64).abs().max().item() < 5e-5
    ll = d.log_likelihood(x_u8.to(DEV), dtype=torch.float64)
    # trained-like data: an element within rounding of the 1e-5 branch threshold may take either branch (see util)
    rel = ((ll.detach().cpu() - ll64).abs() / ll64.abs())
    ok = rel <= LL_RTOL
    assert bool(ok.all()) or dist == "trained" and float(ok.float().mean()) >= 0.5
    (d.log_likelihood(x_u8.to(DEV)) * g_image.to(DEV)).sum().backward()
    if bool(ok.all()):
        assert_grad_close(pd.grad, grad64, M)
    else:
        assert relnorm(pd.grad.cpu()[ok], grad64[ok]) <= GRAD_RTOL


def test_fused_iwae_step_plain(V):
    from vae_mdl_b200 import functional as F
    S, B, H, W, M = 4, 5, 8, 8, 10
    params, x_u8, g = canonical(77, S, B, H, W, M)
    p64 = params.double().requires_grad_(True)
    ll64 = O.mdl_plain_log_prob(p64, O.normalize_u8(x_u8, torch.float64)).sum((-1, -2))
    extra = (ll64.detach().mean(0, keepdim=True) - ll64.detach()) + torch.randn(S, B, generator=g).double()
    loss64 = -O.logmeanexp(ll64 + extra, 0).mean()
    loss64.backward()
    ll, log_w, lme_b, elbo, g_ll = F.modl_iwae_forward(params.to(DEV), x_u8.to(DEV), extra.float().to(DEV), plain=True)
    assert abs(-elbo.item() - loss64.item()) <= LL_RTOL * abs(loss64.item())
    dp = F.modl_backward(params.to(DEV), x_u8.to(DEV), g_image=g_ll, plain=True)
    assert_grad_close(dp, p64.grad, M)


def test_sample_mean_and_attributes(V):
    g = torch.Generator().manual_seed(8)
    B, H, W, M = 3, 8, 8, 5
    p = torch.randn(B, H, W, 10 * M, generator=g)
    n = 3
    um = torch.rand(n, B, H, W, M, generator=g) * (1 - 2e-5) + 1e-5
    ul = torch.rand(n, B, H, W, 3, M, generator=g) * (1 - 2e-5) + 1e-5
    d = V.PixelMixtureDiscretizedLogistic(p.to(DEV))
    x, xq, idx = d.sample(n, u_mix=um.to(DEV), u_log=ul.to(DEV), return_index=True, return_quantised=True)
    want, widx = O.mdl_plain_sample(p.expand(n, B, H, W, 10 * M), um, ul)
    assert x.shape == (n, B, H, W, 3)
    assert int((idx.cpu().long() != widx).sum()) == 0
    assert int((xq.cpu() != O.quantise(want)).sum()) == 0
    assert (x.cpu().double() - want).abs().max().item() < 1e-6
    m = d.mean(u_mix=um[:1].to(DEV))
    wm, _ = O.mdl_plain_sample(p, um[0], None)
    assert m.shape == (B, H, W, 3) and (m.cpu().double() - wm).abs().max().item() < 1e-6
    assert d.sample().shape == (B, H, W, 3) and d.sample([2]).shape == (2, B, H, W, 3)
    assert d.n_mix == M and d.mix_logits.shape == (B, H, W, M) and d.loc.shape == (B, H, W, 3, M)
    loc64, ls64, _ = O.mdl_plain_get_mixture_params(p.double())
    assert relnorm(d.loc, loc64) < 1e-6 and relnorm(d.logscale, ls64) < 1e-6
    with pytest.raises(ValueError):
        V.PixelMixtureDiscretizedLogistic(p.to(DEV), low=1.0, high=1.0)
    with pytest.raises(ValueError):                      # 2-level grid: bin width 2 > 0.049 (include/vaemdl.h)
        V.PixelMixtureDiscretizedLogistic(p.to(DEV), levels=2.0)
    # the class built with its own (low, high, levels) (utils/mdl_plain.py:18): sampler clipped to [low, high], mean to [-1, 1]
    lo_, hi_ = -0.4, 0.7
    d2 = V.PixelMixtureDiscretizedLogistic(p.to(DEV), low=lo_, high=hi_, levels=64.0)
    x2, xq2, idx2 = d2.sample(n, u_mix=um.to(DEV), u_log=ul.to(DEV), return_index=True, return_quantised=True)
    want2, widx2 = O.mdl_plain_sample(p.expand(n, B, H, W, 10 * M), um, ul, lo_, hi_)
    assert int((idx2.cpu().long() != widx2).sum()) == 0 and int((xq2.cpu() != O.quantise(want2)).sum()) == 0
    assert (x2.cpu().double() - want2).abs().max().item() < 1e-6
    assert x2.min().item() >= (lo_ + 1) / 2 - 1e-6 and x2.max().item() <= (hi_ + 1) / 2 + 1e-6
    assert (d2.mean(u_mix=um[:1].to(DEV)).cpu().double() - wm).abs().max().item() < 1e-6


@pytest.mark.parametrize("low,high,levels,M", [(-0.95, 0.9, 64.0, 5), (-1.0, 1.0, 128.0, 10), (-0.5, 0.5, 256.0, 7),
                                               (-1.0, 1.0, 42.0, 3), (-2.0, 3.0, 1024.0, 20)])
def test_plain_mixture_with_its_own_bins(V, low, high, levels, M):
    """PixelMixtureDiscretizedLogistic(parameters, low, high, levels) (utils/mdl_plain.py:18): edge tests x <= low /
    x >= high, bin width (high - low) / (levels - 1) (utils/discretized_logistic.py:18-21, :71-76): per-pixel values,
    per-image sums and the gradient against the float64 oracle, canonical and trained-like (narrow-scale) parameters."""
    from vae_mdl_b200 import functional as F
    S, B, H, W = 3, 4, 16, 16
    for seed, maker in ((1, canonical), (2, trained_like)):
        params, x_u8, g = maker(40 + seed + M, S, B, H, W, M)
        p64 = params.double().requires_grad_(True)
        x64 = O.normalize_u8(x_u8, torch.float64)
        lp64 = O.mdl_plain_log_prob(p64, x64, low, high, levels)
        ll64 = lp64.sum((-1, -2))
        gi = torch.randn(S, B, generator=g)
        (ll64 * gi.double()).sum().backward()
        d = V.PixelMixtureDiscretizedLogistic(params.to(DEV).requires_grad_(True), low=low, high=high, levels=levels)
        lp = d.log_prob(x_u8.to(DEV))
        keep = ~threshold_ambiguous_plain(p64.detach(), x64, low, high, levels)
        assert ((lp.detach().cpu().double() - lp64.detach()).abs()[keep]).max().item() < 2e-4
        ll = d.log_likelihood(x_u8.to(DEV), dtype=torch.float64)
        ok_img = keep.reshape(S, B, -1).all(-1)
        assert ((ll.cpu() - ll64.detach()).abs() / ll64.detach().abs())[ok_img].max().item() <= LL_RTOL
        dp = F.modl_backward(params.to(DEV), x_u8.to(DEV), g_image=gi.to(DEV), plain=(low, high, levels))
        if bool(ok_img.all()):
            assert_grad_close(dp, p64.grad, M)
        else:
            assert relnorm(dp.cpu()[ok_img], p64.grad[ok_img]) <= GRAD_RTOL


def test_golden_plain_mixture_and_latent_terms(V):
    """tests/golden/plain_m5_latent.npz: trained-like parameters (narrow scales, both edges), sampler, mean, and the
    latent-side Normal terms with their gradients."""
    from util import golden
    from vae_mdl_b200 import functional as F
    z = golden("plain_m5_latent")
    M = 5
    params = torch.from_numpy(z["params"]).to(DEV)
    x_u8 = torch.from_numpy(z["x_u8"]).to(DEV)
    d = V.PixelMixtureDiscretizedLogistic(params)
    ll = d.log_likelihood(x_u8, dtype=torch.float64).cpu()
    want = torch.from_numpy(z["ll"])
    assert ((ll - want).abs() / want.abs()).max().item() <= LL_RTOL
    dp = F.modl_backward(params, x_u8, g_image=torch.from_numpy(z["g_image"]).to(DEV), plain=True)
    assert_grad_close(dp, z["grad"], M)
    um, ul = torch.from_numpy(z["u_mix"]).to(DEV), torch.from_numpy(z["u_log"]).to(DEV)
    x, xq, idx = d.sample(1, u_mix=um[None], u_log=ul[None], return_index=True, return_quantised=True)
    assert torch.equal(idx[0].cpu(), torch.from_numpy(z["idx"])) and torch.equal(xq[0].cpu(), torch.from_numpy(z["q_sample"]))
    assert (x[0].cpu().double() - torch.from_numpy(z["x_sample"])).abs().max().item() < 1e-6
    assert (d.mean(u_mix=um[None]).cpu().double() - torch.from_numpy(z["x_mean"])).abs().max().item() < 1e-6
    zz, ql, qs = (torch.from_numpy(z[k]).to(DEV) for k in ("z", "q_loc", "q_scale"))
    terms = [(zz, None, None, 0.7), (zz, ql, qs, -0.7)]
    extra, sums = F.latent_terms(terms)
    assert relnorm(extra, torch.from_numpy(z["extra"])) < 1e-6
    assert relnorm(sums[0], torch.from_numpy(z["lpz"])) < 1e-6 and relnorm(sums[1], torch.from_numpy(z["lqzx"])) < 1e-6
    dz, dloc, dsc = F.latent_terms_backward(terms, torch.from_numpy(z["g_image"]).to(DEV), share_dz=((0, 1),))
    assert relnorm(dz[0], torch.from_numpy(z["dz"])) < 1e-5
    assert relnorm(dloc[1], torch.from_numpy(z["dq_loc"])) < 1e-5 and relnorm(dsc[1], torch.from_numpy(z["dq_scale"])) < 1e-5
