"""GPU parity tests of logmeanexp / the IWAE tail / the loss functions (utils/utils.py:9-11, models/loss.py, model06)."""
import math

import pytest
import torch
import torch.distributions as td

import oracle as O
from util import GRAD_RTOL, LL_RTOL, assert_grad_close, canonical, relnorm

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def V(built_lib):
    import vae_mdl_b200
    return vae_mdl_b200


@pytest.mark.parametrize("shape,axis", [((5, 64), 0), ((5000, 1), 0), ((16, 256), 0), ((7, 3, 5), 1), ((4, 6), -1), ((1, 9), 0)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_logmeanexp_fwd_bwd(V, shape, axis, dtype):
    g = torch.Generator().manual_seed(sum(shape))
    lw = torch.randn(*shape, generator=g) * 6 - 2e4   # realistic magnitude: per-image log-weights ~ -2e4
    lw64 = lw.double().requires_grad_(True)
    w = torch.randn(O.logmeanexp(lw64, axis).shape, generator=g)
    out64 = O.logmeanexp(lw64, axis)
    (out64 * w.double()).sum().backward()
    lwd = lw.to(DEV).to(dtype).requires_grad_(True)
    out = V.logmeanexp(lwd, axis)
    assert out.shape == out64.shape
    assert ((out.detach().cpu().double() - out64.detach()).abs() / out64.detach().abs()).max().item() < 2e-7
    (out * w.to(DEV)).sum().backward()
    assert relnorm(lwd.grad, lw64.grad) < 1e-5


def test_iwae_tail_matches_oracle(V):
    from vae_mdl_b200 import functional as F
    g = torch.Generator().manual_seed(2)
    S, B = 5, 64
    ll = (torch.randn(S, B, generator=g, dtype=torch.float64) * 3 - 2e4)
    extra = torch.randn(S, B, generator=g)
    ll64 = ll.clone().requires_grad_(True)
    loss64, met = O.iwae_loss(ll64.reshape(S, B, 1, 1, 1), extra.double(), torch.zeros(S, B, dtype=torch.float64), (B, 32, 32, 3))
    loss64.backward()
    log_w, lme_b, elbo, g_ll = F.iwae_tail(ll.to(DEV), extra.to(DEV))
    assert abs(-elbo.item() - loss64.item()) <= 1e-6 * abs(loss64.item())
    assert relnorm(g_ll, ll64.grad) < 1e-5
    assert relnorm(log_w, (ll + extra.double())) < 1e-6
    # float32 input: same API, float32-limited accuracy
    _, _, elbo32, g32 = F.iwae_tail(ll.float().to(DEV), extra.to(DEV))
    assert abs(-elbo32.item() - loss64.item()) <= 1e-6 * abs(loss64.item())
    # sharded normaliser: two half-batches with b_total = B add up to the whole
    _, _, e0, g0 = F.iwae_tail(ll[:, :40].to(DEV), extra[:, :40].to(DEV), b_total=B)
    _, _, e1, g1 = F.iwae_tail(ll[:, 40:].to(DEV), extra[:, 40:].to(DEV), b_total=B)
    assert abs((e0 + e1).item() - elbo.item()) <= 1e-6 * abs(elbo.item())
    assert relnorm(torch.cat([g0, g1], 1), ll64.grad) < 1e-5


@pytest.mark.parametrize("S,B,H,W,M,b_total", [(5, 6, 32, 32, 10, 0), (3, 4, 16, 16, 5, 0), (4, 3, 8, 8, 30, 0),
                                                  (2, 5, 8, 8, 20, 0), (7, 9, 4, 4, 7, 0), (1200, 2, 8, 8, 10, 0),
                                                  (6000, 1, 4, 8, 10, 0), (5, 11, 8, 8, 10, 40), (3, 3, 2, 2, 10, 0)])
def test_fused_forward_finish_matches_oracle(V, S, B, H, W, M, b_total):
    """vaemdl_modl_iwae_fwd (forward kernel + ONE finish kernel) against models/loss.py:32-37 in float64, including the
    routes that fall back to the separate tail: any-M kernel (M=7), images smaller than a tile (2x2), more than 512
    importance samples (1200, 6000)."""
    from vae_mdl_b200 import functional as F
    params, x_u8, g = canonical(100 + S + M, S, B, H, W, M)
    p64 = params.double().requires_grad_(True)
    x64 = O.normalize_u8(x_u8, torch.float64)
    lp64 = O.modl_log_prob(p64, x64)
    ll64 = lp64.sum((-1, -2, -3))
    extra = (ll64.detach().mean(0, keepdim=True) - ll64.detach()) + torch.randn(S, B, generator=g).double()
    loss64, _ = O.iwae_loss(lp64, extra, torch.zeros_like(extra), x64.shape)
    if b_total:
        loss64 = loss64 * B / b_total
    g64 = -torch.softmax(ll64.detach() + extra, 0) / (b_total or B)   # d loss / d lpxz (models/loss.py:34-37)
    ll, log_w, lme_b, elbo, g_ll = F.modl_iwae_forward(params.to(DEV), x_u8.to(DEV), extra.float().to(DEV), b_total)
    assert ll.dtype == torch.float64 and ll.shape == (S, B)
    assert ((ll.cpu() - ll64.detach()).abs() / ll64.detach().abs()).max().item() <= LL_RTOL
    assert abs(-elbo.item() - loss64.item()) <= LL_RTOL * abs(loss64.item())
    assert relnorm(log_w, ll64.detach() + extra) < 1e-6
    lme64 = O.logmeanexp(ll64.detach() + extra, 0)
    assert relnorm(lme_b, lme64) < 1e-6
    assert relnorm(g_ll, g64) < GRAD_RTOL
    # the step built on it gives the parameter gradient of the loss
    loss, lpxz, dparams = V.modl_iwae_step(params.to(DEV), x_u8.to(DEV), extra.float().to(DEV), b_total=b_total)
    loss64.backward()
    assert_grad_close(dparams, p64.grad, M)
    # bitwise reproducible (fixed-order reductions) and independent of what ran before on the stream
    ll2, _, _, elbo2, g2 = F.modl_iwae_forward(params.to(DEV), x_u8.to(DEV), extra.float().to(DEV), b_total)
    assert torch.equal(ll, ll2) and torch.equal(elbo, elbo2) and torch.equal(g_ll, g2)


def _setup_model05_like(g, S, B, H, W, M, n_latent=20):
    params = torch.randn(S, B, H, W, 10 * M, generator=g)
    x_u8 = torch.randint(0, 256, (B, H, W, 3), dtype=torch.uint8, generator=g)
    q_loc = torch.randn(B, n_latent, generator=g)
    q_scale = torch.rand(B, n_latent, generator=g) + 0.5
    z = q_loc + q_scale * torch.randn(S, B, n_latent, generator=g)
    return params, x_u8, q_loc, q_scale, z


def test_iwae_loss_drop_in(V):
    """models/loss.py:26-55 with torch.distributions.Normal for pz / qzx and our MoDL class for pxz (model05)."""
    g = torch.Generator().manual_seed(4)
    S, B, H, W, M = 5, 6, 8, 8, 5
    params, x_u8, q_loc, q_scale, z = _setup_model05_like(g, S, B, H, W, M)
    # oracle, float64
    p64 = params.double().requires_grad_(True)
    x64 = O.normalize_u8(x_u8, torch.float64)
    lpz = td.Normal(0.0, 1.0).log_prob(z.double()).sum(-1)
    lqzx = td.Normal(q_loc.double(), q_scale.double()).log_prob(z.double()).sum(-1)
    loss64, met64 = O.iwae_loss(O.modl_log_prob(p64, x64), lpz, lqzx, x64.shape, beta=0.9)
    loss64.backward()
    # product
    pd = params.to(DEV).requires_grad_(True)
    x01 = O.normalize_u8(x_u8).to(DEV)
    zd = z.to(DEV)
    pz = td.Normal(torch.zeros_like(zd), torch.ones_like(zd))
    pz.axes = [-1]
    qzx = td.Normal(q_loc.to(DEV), q_scale.to(DEV))
    qzx.axes = [-1]
    pxz = V.MixtureDiscretizedLogistic(pd)
    loss, met = V.iwae_loss(x01, zd, pz, qzx, pxz, beta=0.9)
    assert set(met) == {"iwae_elbo", "bpd", "lpxz", "lqzx", "lpz", "kl"}
    assert met["lpxz"].shape == (S, B) and met["lpxz"].dtype == torch.float32
    assert abs(loss.item() - loss64.item()) <= LL_RTOL * abs(loss64.item())
    assert abs(met["bpd"].item() - met64["bpd"].item()) <= LL_RTOL * abs(met64["bpd"].item())
    assert relnorm(met["kl"], met64["kl"]) < 1e-5
    loss.backward()
    assert_grad_close(pd.grad, p64.grad, M)
    # the generic route (axes that are not the image axes -> per-pixel log_prob + torch reduce) gives the same loss
    pxz2 = V.MixtureDiscretizedLogisticOpenaiIWAE(params.to(DEV))
    loss2, _ = V.iwae_loss(x01, zd, pz, qzx, pxz2, beta=0.9)
    assert abs(loss2.item() - loss64.item()) <= LL_RTOL * abs(loss64.item())


def test_elbo_loss_and_model06_loss_fn(V):
    g = torch.Generator().manual_seed(5)
    S, B, H, W = 5, 4, 8, 8
    both = torch.randn(S, B, H, W, 6, generator=g)
    x_u8 = torch.randint(0, 256, (B, H, W, 3), dtype=torch.uint8, generator=g)
    x01 = O.normalize_u8(x_u8)
    z1 = torch.randn(S, B, 8, generator=g)
    z2 = torch.randn(S, B, 4, generator=g)
    mk = lambda t: t.to(DEV)  # noqa: E731
    # ---- elbo_loss with the plain DL head (model03 shape)
    lp64 = O.dlogistic_log_prob(x01.double(), both[..., :3].double(), both[..., 3:].double(), 0.0, 1.0, 256.0)
    lpz = td.Normal(0.0, 1.0).log_prob(z1.double()).sum(-1)
    lq = td.Normal(0.1, 1.3).log_prob(z1.double()).sum(-1)
    want, _ = O.elbo_loss(lp64, lpz, lq)
    bd = mk(both)
    mu, lstd = torch.split(bd, 3, dim=-1)
    pxz = V.DiscretizedLogistic(mu, lstd, low=0.0, high=1.0, levels=256.0)
    pxz.axes = [-1, -2, -3]                                  # models/model03.py:143
    pz = td.Normal(torch.zeros_like(mk(z1)), torch.ones_like(mk(z1)))
    pz.axes = [-1]
    qz = td.Normal(torch.full_like(mk(z1), 0.1), torch.full_like(mk(z1), 1.3))
    qz.axes = [-1]
    got, met = V.elbo_loss(mk(x01), mk(z1), pz, qz, pxz)
    assert abs(got.item() - want.item()) <= LL_RTOL * abs(want.item())
    # ---- model06 loss_fn (models/model06.py:38-72)
    lpz2 = td.Normal(0.0, 1.0).log_prob(z2.double()).sum(-1)
    lqz2z1 = td.Normal(0.2, 0.9).log_prob(z2.double()).sum(-1)
    lpz1z2 = td.Normal(-0.1, 1.1).log_prob(z1.double()).sum(-1)
    lqz1x = lq
    want6, met6 = O.model06_loss(lp64, lpz2, lqz2z1, lpz1z2, lqz1x, x01.shape)
    DT = V.DistributionTuple
    pz_top = td.Normal(torch.zeros_like(mk(z2)), torch.ones_like(mk(z2)))
    pz_top.axes = [-1]
    got6, m6 = V.loss_fn(
        mk(x01), pz_top,
        DT(qz, mk(z1), (-1,)),
        DT(td.Normal(torch.full_like(mk(z2), 0.2), torch.full_like(mk(z2), 0.9)), mk(z2), (-1,)),
        DT(td.Normal(torch.full_like(mk(z1), -0.1), torch.full_like(mk(z1), 1.1)), None, (-1,)),
        DT(pxz, None, (-1, -2, -3)))
    assert abs(got6.item() - want6.item()) <= LL_RTOL * abs(want6.item())
    assert abs(m6["bpd"].item() - met6["bpd"].item()) <= LL_RTOL * abs(met6["bpd"].item())
    assert set(m6) == set(met6)


def test_log_sum_exp_helpers(V):
    g = torch.Generator().manual_seed(6)
    x = torch.randn(3, 4, 4, 10, generator=g) * 5
    a = V.log_sum_exp(x.to(DEV)).cpu()
    assert torch.allclose(a, O.log_sum_exp(x), atol=1e-5)
    b = V.log_prob_from_logits(x.to(DEV)).cpu()
    assert torch.allclose(b, O.log_prob_from_logits(x), atol=1e-5)
    assert V.int_shape(x) == [3, 4, 4, 10]


def test_fused_iwae_loss_gradients_wrt_latents_and_parameters(V):
    """models/loss.py:26-46 end to end: the fused objective's gradients w.r.t. the decoder output, z, and the encoder's
    Normal parameters against torch-CPU float64 autograd over the oracle; z itself depends on (q_loc, q_scale) through the
    reparameterisation, as in models/model05.py:124-125."""
    g = torch.Generator().manual_seed(21)
    S, B, H, W, M, D = 5, 6, 8, 8, 10, 20
    params, x_u8, q_loc, q_scale, _ = _setup_model05_like(g, S, B, H, W, M, D)
    eps = torch.randn(S, B, D, generator=g)
    beta = 0.8
    x64 = O.normalize_u8(x_u8, torch.float64)
    pd = params.to(DEV).requires_grad_(True)
    ql = q_loc.to(DEV).requires_grad_(True)
    qs = q_scale.to(DEV).requires_grad_(True)
    z = ql + qs * eps.to(DEV)
    pz = td.Normal(torch.zeros_like(z), torch.ones_like(z))
    pz.axes = [-1]
    qzx = td.Normal(ql, qs)
    qzx.axes = [-1]
    pxz = V.MixtureDiscretizedLogistic(pd)
    loss, met = V.iwae_loss(O.normalize_u8(x_u8).to(DEV), z, pz, qzx, pxz, beta=beta)
    # oracle: the same graph in float64 on the CPU
    p64b = params.double().requires_grad_(True)
    ql64b = q_loc.double().requires_grad_(True)
    qs64b = q_scale.double().requires_grad_(True)
    z64b = ql64b + qs64b * eps.double()
    lossb, metb = O.iwae_loss(O.modl_log_prob(p64b, x64), td.Normal(0.0, 1.0).log_prob(z64b).sum(-1),
                              td.Normal(ql64b, qs64b).log_prob(z64b).sum(-1), x64.shape, beta=beta)
    lossb.backward()
    assert abs(loss.item() - lossb.item()) <= LL_RTOL * abs(lossb.item())
    assert relnorm(met["lpz"], metb["lpz"]) < 1e-6 and relnorm(met["lqzx"], metb["lqzx"]) < 1e-6
    assert relnorm(met["kl"], metb["kl"]) < 1e-5
    loss.backward()
    assert_grad_close(pd.grad, p64b.grad, M)
    assert relnorm(ql.grad, ql64b.grad) < GRAD_RTOL and relnorm(qs.grad, qs64b.grad) < GRAD_RTOL


def test_iwae_loss_with_the_references_constant_prior_and_singleton_sample_axis(V):
    """The reference's own prior is ``Normal(0.0, 1.0)`` built from Python numbers (models/model05.py:108,
    models/model06.py:185): 0-dim CPU tensors.  The fused route takes it as the parameter-free standard-normal term; an
    encoder Normal whose parameters carry a singleton sample axis ``[1,B,D]`` gets its gradient back in that shape."""
    g = torch.Generator().manual_seed(33)
    S, B, H, W, M, D = 4, 5, 8, 8, 5, 12
    params, x_u8, q_loc, q_scale, _ = _setup_model05_like(g, S, B, H, W, M, D)
    eps = torch.randn(S, B, D, generator=g)
    x64 = O.normalize_u8(x_u8, torch.float64)
    pd = params.to(DEV).requires_grad_(True)
    ql = q_loc.to(DEV)[None].clone().requires_grad_(True)     # [1, B, D]
    qs = q_scale.to(DEV)[None].clone().requires_grad_(True)
    z = ql + qs * eps.to(DEV)
    pz = td.Normal(0.0, 1.0)
    pz.axes = [-1]
    qzx = td.Normal(ql, qs)
    qzx.axes = [-1]
    pxz = V.MixtureDiscretizedLogistic(pd)
    from vae_mdl_b200 import loss as Lmod
    assert Lmod._fusable(O.normalize_u8(x_u8).to(DEV), z, pz, qzx, pxz)
    loss, met = V.iwae_loss(O.normalize_u8(x_u8).to(DEV), z, pz, qzx, pxz, beta=0.7)
    p64 = params.double().requires_grad_(True)
    ql64 = q_loc.double()[None].clone().requires_grad_(True)
    qs64 = q_scale.double()[None].clone().requires_grad_(True)
    z64 = ql64 + qs64 * eps.double()
    loss64, met64 = O.iwae_loss(O.modl_log_prob(p64, x64), td.Normal(0.0, 1.0).log_prob(z64).sum(-1),
                                td.Normal(ql64, qs64).log_prob(z64).sum(-1), x64.shape, beta=0.7)
    loss64.backward()
    assert abs(loss.item() - loss64.item()) <= LL_RTOL * abs(loss64.item())
    assert relnorm(met["lpz"], met64["lpz"]) < 1e-6
    loss.backward()
    assert ql.grad.shape == (1, B, D) and qs.grad.shape == (1, B, D)
    assert_grad_close(pd.grad, p64.grad, M)
    assert relnorm(ql.grad, ql64.grad) < GRAD_RTOL and relnorm(qs.grad, qs64.grad) < GRAD_RTOL
    # model06's loss_fn with the same constant prior (models/model06.py:185)
    D2 = 7
    mu = torch.rand(S, B, H, W, 3, generator=g)
    lstd = torch.randn(S, B, H, W, 3, generator=g) - 2.0
    z1 = torch.randn(S, B, D, generator=g)
    z2 = torch.randn(S, B, D2, generator=g)
    a1, b1 = torch.randn(B, D, generator=g), torch.rand(B, D, generator=g) + 0.5
    a2, b2 = torch.randn(S, B, D2, generator=g), torch.rand(S, B, D2, generator=g) + 0.5
    a3, b3 = torch.randn(S, B, D, generator=g), torch.rand(S, B, D, generator=g) + 0.5
    dl = V.DiscretizedLogistic(mu.to(DEV), lstd.to(DEV), low=0.0, high=1.0, levels=256.0)
    T = V.DistributionTuple
    out, met6 = V.loss_fn(O.normalize_u8(x_u8).to(DEV), pz,
                          T(td.Normal(a1.to(DEV), b1.to(DEV)), z1.to(DEV), [-1]),
                          T(td.Normal(a2.to(DEV), b2.to(DEV)), z2.to(DEV), [-1]),
                          T(td.Normal(a3.to(DEV), b3.to(DEV)), None, [-1]),
                          T(dl, None, [-1, -2, -3]))
    lp = O.dlogistic_log_prob(x64, mu.double(), lstd.double(), 0.0, 1.0, 256.0).sum((-1, -2, -3))
    lw = (lp + td.Normal(0.0, 1.0).log_prob(z2.double()).sum(-1) - td.Normal(a2.double(), b2.double()).log_prob(z2.double()).sum(-1)
          + td.Normal(a3.double(), b3.double()).log_prob(z1.double()).sum(-1)
          - td.Normal(a1.double(), b1.double()).log_prob(z1.double()).sum(-1))
    want = -(torch.logsumexp(lw, 0) - math.log(S)).mean()
    assert abs(out.item() - want.item()) <= LL_RTOL * abs(want.item())


def test_latent_terms_model06_shapes(V):
    """models/model06.py:40-47: four Normal terms, two latent layers of different width, per-sample and shared parameters;
    forward sums and every gradient against float64 autograd."""
    from vae_mdl_b200 import functional as F
    g = torch.Generator().manual_seed(22)
    S, B, D1, D2 = 5, 7, 12, 5
    z1 = torch.randn(S, B, D1, generator=g)
    z2 = torch.randn(S, B, D2, generator=g)
    q1_loc, q1_sc = torch.randn(B, D1, generator=g), torch.rand(B, D1, generator=g) + 0.5          # q(z1|x): shared over S
    q2_loc, q2_sc = torch.randn(S, B, D2, generator=g), torch.rand(S, B, D2, generator=g) + 0.5    # q(z2|z1): per sample
    p1_loc, p1_sc = torch.randn(S, B, D1, generator=g), torch.rand(S, B, D1, generator=g) + 0.5    # p(z1|z2): per sample
    gex = torch.randn(S, B, generator=g)
    leaves = [t.double().requires_grad_(True) for t in (z1, z2, q1_loc, q1_sc, q2_loc, q2_sc, p1_loc, p1_sc)]
    a1, a2, b1, b2, c1, c2, d1, d2 = leaves
    lpz2 = td.Normal(0.0, 1.0).log_prob(a2).sum(-1)
    lqz2z1 = td.Normal(c1, c2).log_prob(a2).sum(-1)
    lpz1z2 = td.Normal(d1, d2).log_prob(a1).sum(-1)
    lqz1x = td.Normal(b1, b2).log_prob(a1).sum(-1)
    extra64 = (lpz2 - lqz2z1) + (lpz1z2 - lqz1x)                             # models/model06.py:47
    (extra64 * gex.double()).sum().backward()
    dv = lambda t: t.to(DEV)  # noqa: E731
    terms = [(dv(z2), None, None, 1.0), (dv(z2), dv(q2_loc), dv(q2_sc), -1.0),
             (dv(z1), dv(p1_loc), dv(p1_sc), 1.0), (dv(z1), dv(q1_loc), dv(q1_sc), -1.0)]
    extra, sums = F.latent_terms(terms)
    assert relnorm(extra, extra64) < 1e-6
    for got, want in zip(sums, (lpz2, lqz2z1, lpz1z2, lqz1x)):
        assert relnorm(got, want) < 1e-6
    dz, dloc, dsc = F.latent_terms_backward(terms, dv(gex), share_dz=((0, 1), (2, 3)))
    assert dz[1] is None and dz[3] is None and dloc[0] is None
    assert relnorm(dz[0], a2.grad) < 1e-5 and relnorm(dz[2], a1.grad) < 1e-5
    assert relnorm(dloc[1], c1.grad) < 1e-5 and relnorm(dsc[1], c2.grad) < 1e-5
    assert relnorm(dloc[2], d1.grad) < 1e-5 and relnorm(dsc[2], d2.grad) < 1e-5
    assert relnorm(dloc[3], b1.grad) < 1e-5 and relnorm(dsc[3], b2.grad) < 1e-5


@pytest.mark.parametrize("S,B,H,W,M,plain", [(5, 8, 32, 32, 10, False), (5, 16, 32, 32, 5, False), (3, 4, 8, 8, 10, False),
                                              (2, 2, 16, 16, 20, False), (2, 2, 8, 4, 30, False), (1, 3, 5, 7, 10, False),
                                              (32, 1, 8, 8, 10, False), (4, 3, 9, 9, 5, True), (2, 5, 16, 16, 10, True),
                                              (3, 200, 8, 8, 10, False)])
def test_one_launch_step_equals_three_launch_step(built_lib, monkeypatch, S, B, H, W, M, plain):
    """vaemdl_modl_iwae_step: the cooperative one-launch kernel (forward -> grid barrier -> finish -> grid barrier ->
    backward on the resident tile) gives bit-identical per-image sums, weights and gradients to forward + finish +
    backward as three launches, and both match the float64 oracle."""
    from vae_mdl_b200 import functional as F
    g = torch.Generator().manual_seed(7000 + S + 3 * B + H + M)
    params = torch.randn(S, B, H, W, 10 * M, generator=g)
    x_u8 = torch.randint(0, 256, (B, H, W, 3), dtype=torch.uint8, generator=g)
    x_u8.view(-1)[::13] = 0
    x_u8.view(-1)[5::17] = 255
    p64 = params.double().requires_grad_(True)
    x64 = O.normalize_u8(x_u8, torch.float64)
    lp = O.mdl_plain_log_prob(p64, x64) if plain else O.modl_log_prob(p64, x64)[..., 0]
    ll64 = lp.sum((-1, -2))
    extra = (ll64.detach().mean(0, keepdim=True) - ll64.detach()) + torch.randn(S, B, generator=g).double()
    log_w = ll64 + extra.float().double()
    loss64 = -O.logmeanexp(log_w, 0).mean()
    loss64.backward()
    pd, xd, ed = params.to(DEV), x_u8.to(DEV), extra.float().to(DEV)
    monkeypatch.setenv("VAEMDL_FUSED", "1")
    a = F.modl_iwae_step(pd, xd, ed, plain=plain)
    torch.cuda.synchronize()
    assert a[-1] == 1, "the one-launch kernel was not taken"
    monkeypatch.setenv("VAEMDL_FUSED", "0")
    b = F.modl_iwae_step(pd, xd, ed, plain=plain)
    torch.cuda.synchronize()
    assert b[-1] == 3
    for name, u, v in zip(("ll64", "log_w", "lme_b", "elbo", "g_ll", "dparams"), a[:-1], b[:-1]):
        if name == "elbo":
            assert abs(u.item() - v.item()) <= 1e-6 * abs(v.item())
        elif name == "dparams" and M in (10, 20, 30):
            # the one-launch step scales every gradient in the pass that forms it (per-pixel mixture sums handed over by
            # its forward pass); as separate launches n_mix 10 / 20 (and n_mix 30 below 1.2 M pixel-samples) keep the two-pass
            # gradient kernel: round-off apart
            assert relnorm(u, v) <= 2e-6
        else:
            assert torch.equal(u, v), f"{name} differs between the one-launch and the three-launch step"
    assert ((a[0].cpu() - ll64.detach()).abs() / ll64.detach().abs()).max().item() <= 1e-5
    assert abs(-a[3].item() - loss64.item()) <= 1e-5 * abs(loss64.item())
    rel = ((a[5].cpu().double() - p64.grad).norm() / p64.grad.norm()).item()
    assert rel <= 1e-4, rel
    # sharded batch: b_total > B scales the weights and the loss share
    monkeypatch.setenv("VAEMDL_FUSED", "1")
    c = F.modl_iwae_step(pd, xd, ed, b_total=4 * B, plain=plain)
    assert c[-1] == 1
    assert torch.allclose(c[4] * 4.0, a[4], rtol=1e-6, atol=0) and abs(c[3].item() * 4.0 - a[3].item()) <= 1e-5 * abs(a[3].item())
    assert torch.allclose(c[5] * 4.0, a[5], rtol=1e-5, atol=1e-12)


def test_sample_sharded_step_single_process(built_lib):
    """dist.sample_sharded_iwae_step with the real kernels as ll_fn / bwd_fn (world size 1): equals the fused step; the
    two-rank exchange itself is covered by the gloo test (tests/test_host_logic.py) and tools/nccl_split_s_check.py."""
    from vae_mdl_b200 import dist as vdist, functional as F
    S, B, H, W, M = 6, 3, 8, 8, 10
    g = torch.Generator().manual_seed(606)
    params = torch.randn(S, B, H, W, 10 * M, generator=g).to(DEV)
    x_u8 = torch.randint(0, 256, (B, H, W, 3), dtype=torch.uint8, generator=g).to(DEV)
    ll = F.modl_log_likelihood(params, x_u8, dtype=torch.float64)
    extra = (ll.mean(0, keepdim=True) - ll).float() + torch.randn(S, B, generator=g).to(DEV)
    ref = F.modl_iwae_step(params, x_u8, extra)
    loss, lpxz, dp = vdist.sample_sharded_iwae_step(lambda p, x: F.modl_log_likelihood(p, x, dtype=torch.float64),
                                                    lambda p, x, gi: F.modl_backward(p, x, g_image=gi), params, x_u8, extra, S)
    assert torch.equal(lpxz, ref[0])
    assert abs(loss.item() + ref[3].item()) <= 1e-6 * abs(ref[3].item())
    assert relnorm(dp, ref[5]) <= 1e-5


@pytest.mark.parametrize("S,B,world", [(16, 256, 2), (16, 32, 8), (6, 5, 3), (5000, 3, 4)])
def test_split_sample_kernels_equal_the_unsplit_tail(built_lib, S, B, world):
    """vaemdl_iwae_split_local / _combine (importance samples spread over `world` ranks, the all_gather of the (max, sum-exp)
    pairs simulated by a concatenation in rank order): log-mean-exp, ELBO and the local softmax weights equal the unsplit
    vaemdl_iwae_tail and the float64 formula (utils/utils.py:9-11, models/loss.py:34-37)."""
    from vae_mdl_b200 import _abi, dist as vdist, functional as F
    L = _abi.lib()
    g = torch.Generator().manual_seed(S + B)
    ll = (torch.randn(S, B, generator=g, dtype=torch.float64) * 3 - 2.0e4).to(DEV)
    extra = torch.randn(S, B, generator=g).to(DEV)
    lw = ll + extra.double()
    want_lme = torch.logsumexp(lw, 0) - math.log(S)
    want_g = -torch.softmax(lw, 0) / B
    st = _abi.stream_ptr(torch.device(DEV))
    bounds = [vdist.shard_bounds(S, r, world) for r in range(world)]
    pairs = torch.empty(world, 2, B, dtype=torch.float64, device=DEV)
    for r, (lo, hi) in enumerate(bounds):
        rc = L.vaemdl_iwae_split_local(ll[lo:hi].contiguous().data_ptr(), extra[lo:hi].contiguous().data_ptr(), hi - lo, B,
                                       pairs[r].data_ptr(), st)
        assert rc == 0
    for r, (lo, hi) in enumerate(bounds):
        lme = torch.empty(B, device=DEV)
        elbo = torch.empty(1, device=DEV)
        g_ll = torch.empty(hi - lo, B, device=DEV)
        log_w = torch.empty(hi - lo, B, device=DEV)
        rc = L.vaemdl_iwae_split_combine(ll[lo:hi].contiguous().data_ptr(), extra[lo:hi].contiguous().data_ptr(), hi - lo, B,
                                         pairs.data_ptr(), world, S, 0, log_w.data_ptr(), lme.data_ptr(), elbo.data_ptr(),
                                         g_ll.data_ptr(), st)
        assert rc == 0
        assert ((lme.double() - want_lme).abs() / want_lme.abs()).max().item() < 1e-6
        assert abs(elbo.item() - want_lme.mean().item()) <= 1e-6 * abs(want_lme.mean().item())
        assert relnorm(g_ll, want_g[lo:hi]) < 1e-6
        assert relnorm(log_w, lw[lo:hi]) < 1e-6
        if r == 0:
            first = (lme.clone(), elbo.clone())
        else:   # bit-identical on every rank
            assert torch.equal(lme, first[0]) and torch.equal(elbo, first[1])
    # world size 1 through the host helper == the fused tail
    elbo1, g1 = vdist.split_sample_tail(ll, extra, S)
    _, lme_t, elbo_t, g_t = F.iwae_tail(ll, extra)
    assert abs(elbo1.item() - elbo_t.item()) <= 1e-6 * abs(elbo_t.item()) and relnorm(g1, g_t) < 1e-5


def test_peer_elbo_exchange_single_process(built_lib):
    """The peer-memory exchange of the ELBO shares (include/vaemdl.h: vaemdl_peer_next / vaemdl_peer_elbo_sum) with every
    "rank" mapped into one local buffer: each ELBO-producing route (one-launch step, forward + finish, the unfused tail for
    S > 512, the plain discretized logistic) publishes {seq, share} at [seq % ring][rank]; the reader adds the words of a
    step in rank order and answers NaN for a step that never arrived.  (Two real ranks over NVLink: tools/peer_check.py.)"""
    import ctypes
    from vae_mdl_b200 import _abi, functional as F
    from vae_mdl_b200.peer import ElboExchange
    L = _abi.lib()
    dev = torch.device(DEV)
    n, ring = 3, 4
    buf = torch.zeros(ring * n, dtype=torch.int64, device=dev)
    out = torch.empty(1, device=dev)
    st = _abi.stream_ptr(dev)

    def attach(rank, seq):
        p = _abi.VaemdlPeer()
        for r in range(n):
            p.slots[r] = buf.data_ptr()
        p.n_ranks, p.rank, p.ring, p.seq = n, rank, ring, seq
        assert L.vaemdl_peer_next(ctypes.byref(p)) == 0

    g = torch.Generator().manual_seed(11)
    S, B, H, W, M = 4, 6, 16, 16, 10
    shares = []
    for seq in (1, 2, 5):                                  # (5 % 4 == 1: the ring wraps onto step 1's row)
        want = 0.0
        for rank in range(n):
            params = torch.randn(S, B, H, W, 10 * M, generator=g).to(dev)
            x_u8 = torch.randint(0, 256, (B, H, W, 3), dtype=torch.uint8, generator=g).to(dev)
            attach(rank, seq)
            if rank == 0:      # one cooperative launch
                elbo = F.modl_iwae_step(params, x_u8, None, b_total=n * B)[3]
            elif rank == 1:    # forward + fused finish
                elbo = F.modl_iwae_forward(params, x_u8, None, b_total=n * B)[3]
            else:              # plain discretized logistic step
                both = torch.randn(S, B, H, W, 6, generator=g).to(dev)
                elbo = F.dlogistic_iwae_step(both[..., :3], both[..., 3:], x_u8, None, 0.0, 1.0, 256.0, b_total=n * B)[3]
            want += float(elbo.item())
        assert L.vaemdl_peer_elbo_sum(buf.data_ptr(), n, ring, seq, out.data_ptr(), st) == 0
        assert abs(out.item() - want) <= 1e-6 * abs(want)
        shares.append(want)
    words = buf.cpu().view(ring, n)
    assert int(words[1, 0].item() >> 32) == 5 and int(words[2, 2].item() >> 32) == 2
    # a detached call publishes nothing; an overrun step (1 was overwritten by 5) reads as NaN
    before = buf.clone()
    F.modl_iwae_forward(params, x_u8, None)
    assert torch.equal(buf, before)
    assert L.vaemdl_peer_elbo_sum(buf.data_ptr(), n, ring, 1, out.data_ptr(), st) == 0
    assert math.isnan(out.item())
    # the unfused tail (S > 512) publishes through the one-thread push kernel
    params = torch.randn(600, 1, 8, 8, 10 * M, generator=g).to(dev)
    x1 = torch.randint(0, 256, (1, 8, 8, 3), dtype=torch.uint8, generator=g).to(dev)
    for rank in range(n):
        attach(rank, 7)
        e7 = F.modl_iwae_forward(params, x1, None, b_total=n)[3]
    assert L.vaemdl_peer_elbo_sum(buf.data_ptr(), n, ring, 7, out.data_ptr(), st) == 0
    assert abs(out.item() - 3 * e7.item()) <= 1e-6 * abs(3 * e7.item())
    # the host helper with one rank
    ex = ElboExchange(dev)
    seq = ex.attach()
    e1 = F.modl_iwae_forward(params, x1, None)[3]
    assert abs(ex.read(seq).item() - e1.item()) <= 1e-7 * abs(e1.item())
    ex.close()
