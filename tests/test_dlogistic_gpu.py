"""GPU parity tests of the plain discretized-logistic kernels (utils/discretized_logistic.py)."""
import pytest
import torch

import oracle as O
from util import GRAD_RTOL, assert_ll_close, golden, relnorm

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def V(built_lib):
    import vae_mdl_b200
    return vae_mdl_b200


def test_golden_unsplit_conv_output(V):
    """models/model03.py:88-97: out [S,B,32,32,6] -> split -> DiscretizedLogistic(mu, logstd, low=0, high=1, levels=256)."""
    z = golden("dl_small")
    both = torch.from_numpy(z["both"]).to(DEV).requires_grad_(True)
    mu, lstd = torch.split(both, 3, dim=-1)
    d = V.DiscretizedLogistic(mu, lstd, low=0.0, high=1.0, levels=256.0)
    x_u8 = torch.from_numpy(z["x_u8"]).to(DEV)
    x01 = O.normalize_u8(torch.from_numpy(z["x_u8"])).to(DEV)
    lp = d.log_prob(x01)
    assert lp.shape == mu.shape
    assert (lp.detach().cpu().double() - torch.from_numpy(z["lp"])).abs().max().item() < 2e-5
    ll = d.log_likelihood(x_u8)
    assert_ll_close(ll, z["ll"])
    assert_ll_close(d.log_likelihood(x_u8, dtype=torch.float64), z["ll"], rtol=5e-7)
    (ll * torch.from_numpy(z["g_image"]).to(DEV)).sum().backward()
    assert relnorm(both.grad[..., :3], torch.from_numpy(z["dloc"])) < GRAD_RTOL
    assert relnorm(both.grad[..., 3:], torch.from_numpy(z["dls"])) < GRAD_RTOL


@pytest.mark.parametrize("low,high,levels", [(-1.0, 1.0, 256.0), (0.0, 1.0, 256.0), (0.0, 1.0, 32.0)])
def test_separate_tensors_and_ranges(V, low, high, levels):
    g = torch.Generator().manual_seed(int(levels) + int(10 * low))
    S, B, H, W = 2, 3, 8, 8
    loc = (torch.rand(S, B, H, W, 3, generator=g) * (high - low) + low)
    ls = torch.randn(S, B, H, W, 3, generator=g) - 1.0
    k = torch.randint(0, int(levels), (B, H, W, 3), generator=g)
    x = (low + (high - low) * k.double() / (levels - 1)).float()
    w = torch.randn(S, B, H, W, 3, generator=g)
    loc64 = loc.double().requires_grad_(True)
    ls64 = ls.double().requires_grad_(True)
    lp64 = O.dlogistic_log_prob(x.double(), loc64, ls64, low, high, levels)
    (lp64 * w.double()).sum().backward()
    locd = loc.to(DEV).requires_grad_(True)
    lsd = ls.to(DEV).requires_grad_(True)
    d = V.DiscretizedLogistic(locd, lsd, low=low, high=high, levels=levels)
    lp = d.log_prob(x.to(DEV))
    assert (lp.detach().cpu().double() - lp64.detach()).abs().max().item() < 5e-5
    (lp * w.to(DEV)).sum().backward()
    assert relnorm(locd.grad, loc64.grad) < GRAD_RTOL
    assert relnorm(lsd.grad, ls64.grad) < GRAD_RTOL


def test_tests_test_hierarchical_setup_shapes(V):
    """tests/test_hierarchical_setup.py:66-75: DiscretizedLogistic(rand, exp(randn)) on [5,16,32,32,3], x in [0,1]."""
    g = torch.Generator().manual_seed(3)
    loc = torch.rand(5, 16, 32, 32, 3, generator=g)
    ls = torch.exp(torch.randn(5, 16, 32, 32, 3, generator=g))
    x = torch.floor(torch.rand(16, 32, 32, 3, generator=g) * 256) / 255
    d = V.DiscretizedLogistic(loc.to(DEV), ls.to(DEV), low=0.0, high=1.0, levels=256.0)
    ll = d.log_likelihood(x.to(DEV))
    want = O.dlogistic_log_prob(x.double(), loc.double(), ls.double(), 0.0, 1.0, 256.0).sum((-1, -2, -3))
    assert_ll_close(ll, want)


def test_extreme_logscales_match_reference_forward(V):
    """No clamp in this class (utils/discretized_logistic.py:38): vanishing and huge scales."""
    loc = torch.tensor([0.5, 0.5, 0.25, 0.5, 0.5])
    ls = torch.tensor([-100.0, 95.0, -30.0, 20.0, -8.0])
    x = torch.tensor([0.5, 0.5, 128 / 255, 1.0, 0.0])
    want = O.dlogistic_log_prob(x.double(), loc.double(), ls.double(), 0.0, 1.0, 256.0)
    d = V.DiscretizedLogistic(loc.to(DEV), ls.to(DEV), low=0.0, high=1.0, levels=256.0)
    got = d.log_prob(x.to(DEV)).cpu().double()
    for a, b in zip(got.tolist(), want.tolist()):
        if b == float("-inf"):
            assert a == b or a < -1e30
        else:
            assert abs(a - b) <= 1e-5 * max(1.0, abs(b)), (a, b)


def test_sample_explicit_noise(V):
    g = torch.Generator().manual_seed(5)
    loc = torch.randn(4, 8, 8, 3, generator=g) * 0.5
    ls = torch.randn(4, 8, 8, 3, generator=g) - 2
    u = torch.rand(3, 4, 8, 8, 3, generator=g) * (1 - 2e-5) + 1e-5
    d = V.DiscretizedLogistic(loc.to(DEV), ls.to(DEV))
    out = d.sample(3, u=u.to(DEV))
    want = O.dlogistic_sample(loc, ls, u, -1.0, 1.0)
    assert out.shape == (3, 4, 8, 8, 3)
    assert (out.cpu().double() - want).abs().max().item() < 1e-6
    assert d.sample().shape == (4, 8, 8, 3) and d.sample([2]).shape == (2, 4, 8, 8, 3)
    assert d.mean() is d.loc
