"""CPU tests of the oracle itself (no GPU, no CUDA library).

The reference holds no golden vectors for this path ("parity unpinned"), so the oracle is validated by
(1) the two independent formulations in the reference agreeing with each other, (2) an mpmath arbitrary-precision
evaluation of the closed form on hand-built cases that hit every branch, (3) analytic derivatives vs autograd,
(4) the committed golden fixtures (regression pin of the oracle).
"""
import math
import os

import mpmath as mp
import numpy as np
import pytest
import torch

import oracle as O

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
mp.mp.dps = 50


def canonical(seed, S, B, H, W, M, dtype=torch.float64):
    """The reference demos' synthetic distribution (utils/mdl.py:275-296): binned x, randn parameters."""
    g = torch.Generator().manual_seed(seed)
    x_u8 = torch.randint(0, 256, (B, H, W, 3), dtype=torch.uint8, generator=g)
    params = torch.randn(S, B, H, W, 10 * M, generator=g).to(dtype)
    return params, x_u8


# --------------------------------------------------------------------------------------------------
# (1) utils/mdl.py form == utils/mdl_openai.py form on binned data
# --------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("M", [1, 5, 10])
def test_mdl_equals_openai_float64(M):
    params, x_u8 = canonical(1, 3, 4, 8, 8, M)
    x = O.normalize_u8(x_u8, torch.float64)
    a = O.modl_log_prob(params, x)
    b = O.modl_openai_iwae_log_prob(params, x)
    assert a.shape == (3, 4, 8, 8, 1) and b.shape == a.shape
    assert (a - b).abs().max().item() < 1e-11
    # the 4-D wrapper (x in [-1,1], no trailing dim)
    c = O.modl_openai_log_prob(params[0], x * 2 - 1)
    assert c.shape == (4, 8, 8)
    assert (c - a[0, ..., 0]).abs().max().item() < 1e-11


def test_mdl_equals_openai_float32_per_image():
    params, x_u8 = canonical(2, 3, 4, 16, 16, 10, torch.float32)
    x = O.normalize_u8(x_u8)
    a = O.modl_log_prob(params, x).sum((-1, -2, -3))
    b = O.modl_openai_iwae_log_prob(params, x).sum((-1, -2, -3))
    assert ((a - b).abs() / a.abs()).max().item() < 2e-6


def test_edge_conventions_coincide_on_binned_data():
    """2*(0/255)-1 == -1 and 2*(255/255)-1 == 1 exactly in float32; the next bins are inside +-0.999."""
    k = torch.arange(256, dtype=torch.uint8)
    x = O.normalize_u8(k) * 2.0 - 1.0
    assert x[0].item() == -1.0 and x[255].item() == 1.0
    assert ((x <= -1.0) == (x < -0.999)).all() and ((x >= 1.0) == (x > 0.999)).all()


def test_broadcast_x_without_batch_dim():
    """models/model05.py:173 passes x [H,W,3] against parameters [S,1,H,W,10M]."""
    params, x_u8 = canonical(3, 4, 1, 8, 8, 5)
    x = O.normalize_u8(x_u8, torch.float64)
    a = O.modl_log_prob(params, x[0])
    b = O.modl_log_prob(params, x)
    assert a.shape == (4, 1, 8, 8, 1)
    assert torch.equal(a, b)


# --------------------------------------------------------------------------------------------------
# (2) mpmath known answers for every branch
# --------------------------------------------------------------------------------------------------
def mp_sigmoid(v):
    return 1 / (1 + mp.e ** (-v))


def mp_softplus(v):
    return mp.log(1 + mp.e ** v)


def mp_subpixel(x, loc, ls, dx, low, high, width):
    """utils/mdl.py:165-207 in 50-digit arithmetic."""
    x, loc, ls = mp.mpf(x), mp.mpf(loc), mp.mpf(ls)
    inv = mp.e ** (-ls)
    start, stop = (x - loc - dx) * inv, (x - loc + dx) * inv
    if x >= high:
        return -mp_softplus(start)
    if x <= low:
        return stop - mp_softplus(stop)
    prob = mp_sigmoid(stop) - mp_sigmoid(start)
    if prob > mp.mpf("1e-5"):
        return mp.log(prob)
    a = (x - loc) * inv
    return -a - ls - 2 * mp_softplus(-a) + mp.log(width)


def mp_modl_pixel(row, x01, M):
    """One pixel of utils/mdl.py:56-92."""
    row = [mp.mpf(float(v)) for v in row]
    x = [2 * mp.mpf(float(v)) - 1 for v in x01]
    terms = []
    for m in range(M):
        mu = [row[M + c * 3 * M + m] for c in range(3)]
        ls = [max(row[M + c * 3 * M + M + m], mp.mpf(-7)) for c in range(3)]
        k = [mp.tanh(row[M + c * 3 * M + 2 * M + m]) for c in range(3)]
        loc = [mu[0], mu[1] + k[0] * x[0], mu[2] + k[1] * x[0] + k[2] * x[1]]
        dx = mp.mpf(1) / 255
        t = sum(mp_subpixel(x[c], loc[c], ls[c], dx, -1, 1, 2 * dx) for c in range(3))
        terms.append(t + row[m])
    lse = lambda v: max(v) + mp.log(sum(mp.e ** (u - max(v)) for u in v))  # noqa: E731
    return lse(terms) - lse(row[:M])


BRANCH_CASES = [
    # (description, x_u8 triple, overrides applied to every mixture: (mu, raw log-scale, raw coefficient))
    ("normal", (100, 150, 200), (0.0, -1.0, 0.3)),
    ("left-edge", (0, 0, 0), (-0.5, -2.0, 0.5)),
    ("right-edge", (255, 255, 255), (0.5, -2.0, -0.5)),
    ("left-edge, loc beyond the edge", (0, 255, 0), (-1.4, -3.0, 1.5)),
    ("low-prob (narrow, far)", (10, 240, 128), (0.9, -5.0, 0.1)),
    ("clamped log-scale (-9 -> -7)", (128, 127, 129), (0.003, -9.0, 0.0)),
    ("saturated tanh", (30, 200, 90), (0.1, -1.5, 12.0)),
    ("wide scale", (77, 3, 251), (0.0, 3.0, -0.7)),
    ("extreme tail (underflows float32 linear domain)", (0, 128, 255), (3.0, -7.0, 0.0)),
]


@pytest.mark.parametrize("desc,xs,ov", BRANCH_CASES, ids=[c[0] for c in BRANCH_CASES])
def test_modl_known_answers_mpmath(desc, xs, ov):
    M = 3
    g = torch.Generator().manual_seed(7)
    row = torch.randn(10 * M, generator=g, dtype=torch.float64) * 0.1
    for c in range(3):
        row[M + c * 3 * M: M + c * 3 * M + M] += ov[0]
        row[M + c * 3 * M + M: M + c * 3 * M + 2 * M] += ov[1]
        row[M + c * 3 * M + 2 * M: M + c * 3 * M + 3 * M] += ov[2]
    x_u8 = torch.tensor(xs, dtype=torch.uint8)
    x01 = O.normalize_u8(x_u8, torch.float64)
    got = O.modl_log_prob(row.reshape(1, 1, 1, 10 * M), x01.reshape(1, 1, 1, 3)).item()
    want = mp_modl_pixel(row.tolist(), x01.tolist(), M)
    assert abs(got - float(want)) <= 1e-10 * max(1.0, abs(float(want))), (desc, got, float(want))


@pytest.mark.parametrize("x,loc,ls", [(0.3, 0.1, -1.0), (0.0, 0.4, -2.0), (1.0, 0.7, -3.0), (0.5, 0.9, -6.0), (0.2, 0.2, 2.0)])
def test_dlogistic_known_answers_mpmath(x, loc, ls):
    width = mp.mpf(1) / 255
    want = mp_subpixel(x, loc, ls, width / 2, 0, 1, width)
    got = O.dlogistic_log_prob(torch.tensor(x, dtype=torch.float64), torch.tensor(loc, dtype=torch.float64),
                               torch.tensor(ls, dtype=torch.float64), 0.0, 1.0, 256.0).item()
    assert abs(got - float(want)) <= 1e-10 * max(1.0, abs(float(want)))


def test_branch_frequencies_canonical_distribution():
    """left 0.4 %, right 0.4 %, low-probability 2.6 % of (pixel, channel, mixture) elements."""
    params, x_u8 = canonical(11, 2, 8, 32, 32, 10)
    x = O.normalize_u8(x_u8, torch.float64) * 2 - 1
    loc, logscale, _ = O.ref._mdl_autoregressive_params(params, x)
    inv = torch.exp(-logscale)
    xx = x[..., None]
    prob = torch.sigmoid((xx - loc + 1 / 255) * inv) - torch.sigmoid((xx - loc - 1 / 255) * inv)
    left = (xx <= -1).expand_as(prob)
    right = (xx >= 1).expand_as(prob)
    low = (prob <= 1e-5) & ~left & ~right
    assert 0.002 < left.double().mean().item() < 0.006
    assert 0.002 < right.double().mean().item() < 0.006
    assert 0.015 < low.double().mean().item() < 0.04


# --------------------------------------------------------------------------------------------------
# (3) gradients
# --------------------------------------------------------------------------------------------------
def test_modl_autograd_matches_finite_differences():
    params, x_u8 = canonical(5, 1, 1, 2, 2, 3)
    x = O.normalize_u8(x_u8, torch.float64)
    params = params.requires_grad_(True)
    assert torch.autograd.gradcheck(lambda p: O.modl_log_prob(p, x).sum(), (params,), eps=1e-6, atol=1e-6, rtol=1e-5)


def test_logscale_clamp_gradient_is_zero_below_minus7_and_routed_on_tie():
    M = 2
    row = torch.zeros(1, 1, 1, 10 * M, dtype=torch.float64)
    row[..., M + M] = -9.0      # sR of mixture 0: clamped -> no gradient
    row[..., M + M + 1] = -7.0  # sR of mixture 1: tie -> gradient flows (tf.maximum routes to x when x >= y)
    row = row.requires_grad_(True)
    x = torch.tensor([[[[0.3, 0.5, 0.7]]]], dtype=torch.float64)
    O.modl_log_prob(row, x).sum().backward()
    assert row.grad[0, 0, 0, M + M].item() == 0.0
    assert row.grad[0, 0, 0, M + M + 1].item() != 0.0


def test_logmeanexp_gradient_is_softmax():
    g = torch.Generator().manual_seed(3)
    lw = (torch.randn(6, 5, generator=g, dtype=torch.float64) * 4).requires_grad_(True)
    O.logmeanexp(lw, 0).sum().backward()
    assert torch.allclose(lw.grad, torch.softmax(lw.detach(), 0), atol=1e-14)
    ref = torch.logsumexp(lw.detach(), 0) - math.log(6)
    assert torch.allclose(O.logmeanexp(lw.detach(), 0), ref, atol=1e-12)


def test_iwae_loss_upstream_gradient():
    """d loss / d lpxz[s,b] = -softmax_s(log_w)[s,b] / B  (SURVEY 8a-12)."""
    g = torch.Generator().manual_seed(4)
    S, B = 5, 3
    lpxz_elem = torch.randn(S, B, 2, 2, 1, generator=g, dtype=torch.float64).requires_grad_(True)
    lpz = torch.randn(S, B, generator=g, dtype=torch.float64)
    lqzx = torch.randn(S, B, generator=g, dtype=torch.float64)
    loss, met = O.iwae_loss(lpxz_elem, lpz, lqzx, (B, 2, 2, 3), beta=0.7)
    loss.backward()
    log_w = lpxz_elem.detach().sum((-1, -2, -3)) + 0.7 * (lpz - lqzx)
    want = -torch.softmax(log_w, 0) / B
    assert torch.allclose(lpxz_elem.grad[:, :, 0, 0, 0], want, atol=1e-14)
    assert set(met) == {"iwae_elbo", "bpd", "lpxz", "lqzx", "lpz", "kl"}
    assert math.isclose(met["bpd"].item(), -met["iwae_elbo"].item() / (math.log(2.0) * 12), rel_tol=1e-12)


def test_ref32_close_to_ref64_per_image():
    """per-image LL of the float32 restatement is within ~1e-6 relative of float64."""
    params, x_u8 = canonical(6, 2, 4, 32, 32, 10)
    ll64 = O.modl_log_prob(params, O.normalize_u8(x_u8, torch.float64)).sum((-1, -2, -3))
    ll32 = O.modl_log_prob(params.float(), O.normalize_u8(x_u8)).sum((-1, -2, -3))
    assert ((ll32.double() - ll64).abs() / ll64.abs()).max().item() < 5e-6


# --------------------------------------------------------------------------------------------------
# samplers
# --------------------------------------------------------------------------------------------------
def test_sampler_variants_agree_on_the_selected_column():
    """utils/mdl.py:209-252 (draw for every mixture, then select) == utils/mdl_openai.py:160-193 (select, then draw)
    when the per-mixture noise of the selected component equals the single draw."""
    g = torch.Generator().manual_seed(8)
    N, H, W, M = 3, 4, 4, 5
    l = torch.randn(N, H, W, 10 * M, generator=g)
    u_mix = torch.rand(N, H, W, M, generator=g) * (1 - 2e-5) + 1e-5
    u_all = torch.rand(N, H, W, 3, M, generator=g) * (1 - 2e-5) + 1e-5
    x01, idx = O.modl_sample_mdl(l, u_mix, u_all)
    u_sel = torch.gather(u_all, -1, idx[..., None, None].expand(N, H, W, 3, 1))[..., 0]
    x, idx2 = O.sample_from_discretized_mix_logistic(l, M, u_mix, u_sel)
    assert torch.equal(idx, idx2)
    assert (x * 0.5 + 0.5 - x01).abs().max().item() < 1e-15


def test_gumbel_argmax_follows_softmax():
    g = torch.Generator().manual_seed(9)
    logits = torch.tensor([0.0, 1.0, -1.0, 2.0])
    u = torch.rand(200000, 4, generator=g) * (1 - 2e-5) + 1e-5
    idx = O.gumbel_argmax(logits.expand(200000, 4), u)
    freq = torch.bincount(idx, minlength=4).double() / 200000
    assert (freq - torch.softmax(logits.double(), 0)).abs().max().item() < 5e-3


def test_sample_shapes_and_ranges():
    g = torch.Generator().manual_seed(10)
    S, B, H, W, M = 2, 3, 4, 4, 5
    l = torch.randn(S, B, H, W, 10 * M, generator=g)
    n = 4
    u_mix = torch.rand(n * S * B, H, W, M, generator=g) * (1 - 2e-5) + 1e-5
    u_log = torch.rand(n * S * B, H, W, 3, generator=g) * (1 - 2e-5) + 1e-5
    x, idx = O.modl_openai_iwae_sample(l, n, u_mix, u_log)
    assert x.shape == (n, S, B, H, W, 3) and idx.shape == (n, S, B, H, W)
    assert x.min().item() >= 0.0 and x.max().item() <= 1.0
    x2, _ = O.modl_openai_sample(l[0], n, u_mix[: n * B], u_log[: n * B])
    assert x2.shape == (n, B, H, W, 3) and x2.min().item() >= -1.0 and x2.max().item() <= 1.0
    q = O.quantise(x)
    assert q.dtype == torch.uint8


def test_dlogistic_sample_clip():
    u = torch.tensor([1e-5, 0.5, 1 - 1e-5], dtype=torch.float64)
    out = O.dlogistic_sample(torch.zeros(3), torch.zeros(3), u, low=-1.0, high=1.0)
    assert out[0].item() == -1.0 and out[1].item() == 0.0 and out[2].item() == 1.0


# --------------------------------------------------------------------------------------------------
# (4) golden fixtures pin the oracle
# --------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["modl_m10_randn", "modl_m5_trained", "modl_m30_randn", "modl_m7_ragged"])
def test_oracle_reproduces_modl_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    params = torch.from_numpy(z["params"]).double().requires_grad_(True)
    x64 = O.normalize_u8(torch.from_numpy(z["x_u8"]), torch.float64)
    lp = O.modl_log_prob(params, x64)
    assert np.allclose(lp.detach().numpy()[..., 0], z["lp"], rtol=0, atol=1e-10)
    (lp.sum((-1, -2, -3)) * torch.from_numpy(z["g_image"]).double()).sum().backward()
    assert np.allclose(params.grad.numpy(), z["grad_fixed"], rtol=1e-9, atol=1e-12)


def test_oracle_reproduces_dl_and_sample_golden():
    z = np.load(os.path.join(GOLDEN, "dl_small.npz"))
    both = torch.from_numpy(z["both"]).double()
    lp = O.dlogistic_log_prob(O.normalize_u8(torch.from_numpy(z["x_u8"]), torch.float64), both[..., :3], both[..., 3:],
                              0.0, 1.0, 256.0)
    assert np.allclose(lp.numpy(), z["lp"], rtol=0, atol=1e-10)
    z = np.load(os.path.join(GOLDEN, "sample_m10.npz"))
    x, idx = O.sample_from_discretized_mix_logistic(torch.from_numpy(z["l"]), 10, torch.from_numpy(z["u_mix"]),
                                                    torch.from_numpy(z["u_log"]))
    assert np.array_equal(idx.numpy().astype(np.uint8), z["idx"])
    assert np.array_equal(O.quantise(x * 0.5 + 0.5).numpy(), z["q_openai"])


# --------------------------------------------------------------------------------------------------
# utils/mdl_plain.py (no conditioning on the observed x)
# --------------------------------------------------------------------------------------------------
def test_mdl_plain_equals_conditioned_class_when_the_chain_is_cut():
    """With zero coefficients (tanh(0) = 0) both classes are the same mixture of independent channels."""
    g = torch.Generator().manual_seed(3)
    p = torch.randn(2, 3, 4, 4, 50, generator=g, dtype=torch.float64)
    for c in range(3):
        p[..., 5 + 15 * c + 10:5 + 15 * c + 15] = 0.0
    x = torch.floor(torch.rand(3, 4, 4, 3, generator=g, dtype=torch.float64) * 256) / 255
    assert torch.allclose(O.mdl_plain_log_prob(p, x), O.modl_log_prob(p, x)[..., 0], atol=1e-12)


def test_mdl_plain_chain_follows_the_means_not_x():
    """utils/mdl_plain.py:160-162: a hand-computed single-component pixel."""
    M = 1
    row = torch.tensor([0.3, 0.2, -1.0, 0.5, -0.1, -1.5, -0.4, 0.05, -2.0, 0.7], dtype=torch.float64)  # [logit|muR sR kR|muG sG kG|muB sB kB]
    loc, ls, _ = O.mdl_plain_get_mixture_params(row.view(1, 1, 1, 10 * M))
    k = torch.tanh(torch.tensor([0.5, -0.4, 0.7], dtype=torch.float64))
    want_g = -0.1 + k[0] * 0.2
    want_b = 0.05 + k[1] * 0.2 + k[2] * want_g
    assert torch.allclose(loc.flatten(), torch.stack([torch.tensor(0.2, dtype=torch.float64), want_g, want_b]), atol=1e-15)
    x = torch.tensor([[[[10, 200, 255]]]], dtype=torch.float64) / 255
    lp = O.mdl_plain_log_prob(row.view(1, 1, 1, 10), x)
    want = O.dlogistic_log_prob((x * 2 - 1)[..., None], loc, ls, -1.0, 1.0, 256.0).sum()
    assert abs(lp.item() - want.item()) < 1e-12


def test_mdl_plain_autograd_matches_finite_differences():
    g = torch.Generator().manual_seed(5)
    p = (torch.randn(1, 2, 2, 30, generator=g, dtype=torch.float64) * 0.7).requires_grad_(True)
    x = torch.tensor([[[[3, 128, 255], [0, 17, 99]], [[250, 1, 64], [128, 128, 128]]]], dtype=torch.float64) / 255
    assert torch.autograd.gradcheck(lambda q: O.mdl_plain_log_prob(q, x).sum(), (p,), eps=1e-6, atol=1e-6, rtol=1e-4)


def test_mdl_plain_sample_and_mean():
    g = torch.Generator().manual_seed(6)
    p = torch.randn(2, 4, 4, 50, generator=g)
    um = torch.rand(2, 4, 4, 5, generator=g) * (1 - 2e-5) + 1e-5
    ul = torch.rand(2, 4, 4, 3, 5, generator=g) * (1 - 2e-5) + 1e-5
    x, idx = O.mdl_plain_sample(p, um, ul)
    assert x.shape == (2, 4, 4, 3) and x.min() >= 0 and x.max() <= 1
    m, idx2 = O.mdl_plain_sample(p, um, None)
    assert torch.equal(idx, idx2)
    loc, _, _ = O.mdl_plain_get_mixture_params(p.double())
    want = (torch.clamp(torch.gather(loc, -1, idx[..., None, None].expand(2, 4, 4, 3, 1))[..., 0], -1, 1) + 1) / 2
    assert torch.allclose(m, want, atol=1e-15)
    # u = 0.5 is a zero logistic draw: the sample equals the mean
    x0, _ = O.mdl_plain_sample(p, um, torch.full_like(ul, 0.5))
    assert torch.allclose(x0, m, atol=1e-15)


def test_oracle_reproduces_plain_and_latent_golden():
    import torch.distributions as td
    z = np.load(os.path.join(GOLDEN, "plain_m5_latent.npz"))
    params = torch.from_numpy(z["params"]).double().requires_grad_(True)
    lp = O.mdl_plain_log_prob(params, O.normalize_u8(torch.from_numpy(z["x_u8"]), torch.float64))
    assert np.allclose(lp.detach().numpy(), z["lp"], rtol=0, atol=1e-10)
    (lp.sum((-1, -2)) * torch.from_numpy(z["g_image"]).double()).sum().backward()
    assert np.allclose(params.grad.numpy(), z["grad"], rtol=1e-9, atol=1e-12)
    xs, idx = O.mdl_plain_sample(torch.from_numpy(z["params"]), torch.from_numpy(z["u_mix"]), torch.from_numpy(z["u_log"]))
    assert np.array_equal(idx.numpy().astype(np.uint8), z["idx"]) and np.array_equal(O.quantise(xs).numpy(), z["q_sample"])
    zz = torch.from_numpy(z["z"]).double()
    lpz = td.Normal(0.0, 1.0).log_prob(zz).sum(-1)
    lqzx = td.Normal(torch.from_numpy(z["q_loc"]).double(), torch.from_numpy(z["q_scale"]).double()).log_prob(zz).sum(-1)
    assert np.allclose((0.7 * (lpz - lqzx)).numpy(), z["extra"], rtol=0, atol=1e-10)
