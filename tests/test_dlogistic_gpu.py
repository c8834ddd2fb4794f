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


@pytest.mark.parametrize("S,B,H,W,split,b_total", [(5, 8, 32, 32, True, 0), (3, 5, 8, 8, False, 0), (16, 3, 4, 4, True, 12),
                                                    (700, 2, 4, 4, False, 0), (4, 3, 2, 2, True, 0),
                                                    (2, 3, 6, 6, True, 0), (2, 3, 10, 10, True, 0), (2, 3, 10, 10, False, 0),
                                                    (3, 5, 3, 3, True, 0), (3, 2, 5, 7, False, 0), (2, 7, 9, 8, True, 0)])
def test_fused_iwae_step_matches_oracle(V, S, B, H, W, split, b_total):
    """vaemdl_dlogistic_iwae_fwd + vaemdl_dlogistic_bwd against models/loss.py:32-37 on the model03 head, float64 oracle."""
    g = torch.Generator().manual_seed(S * 7 + B)
    both = torch.randn(S, B, H, W, 6, generator=g)
    both[..., :3] = torch.rand(S, B, H, W, 3, generator=g)
    both[..., 3:] -= 1.5
    x_u8 = torch.randint(0, 256, (B, H, W, 3), dtype=torch.uint8, generator=g)
    b64 = both.double().requires_grad_(True)
    x64 = O.normalize_u8(x_u8, torch.float64)
    lp64 = O.dlogistic_log_prob(x64, b64[..., :3], b64[..., 3:], 0.0, 1.0, 256.0)
    ll64 = lp64.sum((-1, -2, -3))
    extra = (ll64.detach().mean(0, keepdim=True) - ll64.detach()) + torch.randn(S, B, generator=g).double()
    loss64 = -O.logmeanexp(ll64 + extra, 0).sum() / (b_total or B)
    loss64.backward()
    bd = both.to(DEV)
    if split:
        loc, ls = torch.split(bd, 3, dim=-1)          # read in place (ld = 6)
    else:
        loc, ls = bd[..., :3].contiguous(), bd[..., 3:].contiguous()
    loss, lpxz, dloc, dls = V.dlogistic_iwae_step(loc, ls, x_u8.to(DEV), extra.float().to(DEV), 0.0, 1.0, 256.0,
                                                   b_total=b_total)
    assert_ll_close(lpxz, ll64.detach())
    assert abs(loss.item() - loss64.item()) <= 1e-5 * abs(loss64.item())
    assert relnorm(dloc, b64.grad[..., :3]) < GRAD_RTOL
    assert relnorm(dls, b64.grad[..., 3:]) < GRAD_RTOL
    if split:
        assert dls.data_ptr() == dloc.data_ptr() + 12  # one [..,6] gradient buffer, like the un-split input


@pytest.mark.parametrize("shape", [(2, 3, 6, 6, 3), (3, 10, 10, 3), (2, 2, 3, 3, 3), (4, 5, 7, 3), (2, 9, 8, 3), (3, 4, 4, 1), (5, 6)])
@pytest.mark.parametrize("split", [True, False])
@pytest.mark.parametrize("x_u8", [True, False])
def test_elementwise_log_prob_and_grad_every_layout(V, shape, split, x_u8):
    """Element-wise log_prob + autograd through every kernel variant: pixel pairs on the un-split [..,6] layout, pixel
    pairs on separate tensors, and the generic kernel (odd pixel counts, C != 3)."""
    g = torch.Generator().manual_seed(sum(shape) + 2 * split + x_u8)
    C = shape[-1]
    both = torch.randn(*shape[:-1], 2 * C, generator=g)
    both[..., :C] = torch.rand(*shape, generator=g)
    both[..., C:] -= 2.0
    k = torch.randint(0, 256, shape[1:] if len(shape) > 2 else shape, dtype=torch.uint8, generator=g)
    x01 = O.normalize_u8(k)
    w = torch.randn(*shape, generator=g)
    b64 = both.double().requires_grad_(True)
    lp64 = O.dlogistic_log_prob(x01.double(), b64[..., :C], b64[..., C:], 0.0, 1.0, 256.0)
    (lp64 * w.double()).sum().backward()
    bd = both.to(DEV).requires_grad_(True)
    if split:
        loc, ls = torch.split(bd, C, dim=-1)
    else:
        loc, ls = bd[..., :C].contiguous(), bd[..., C:].contiguous()
    d = V.DiscretizedLogistic(loc, ls, low=0.0, high=1.0, levels=256.0)
    lp = d.log_prob(k.to(DEV) if x_u8 else x01.to(DEV))
    assert lp.shape == lp64.shape
    assert (lp.detach().cpu().double() - lp64.detach()).abs().max().item() < 5e-5
    (lp * w.to(DEV)).sum().backward()
    assert relnorm(bd.grad[..., :C], b64.grad[..., :C]) < GRAD_RTOL
    assert relnorm(bd.grad[..., C:], b64.grad[..., C:]) < GRAD_RTOL


@pytest.mark.parametrize("S,B,H,W,interleaved,u8", [(5, 16, 32, 32, True, True), (5, 16, 32, 32, False, False), (3, 4, 8, 8, True, False),
                                                   (2, 3, 16, 12, False, True), (32, 2, 8, 8, True, True), (1, 1, 8, 8, True, True),
                                                   (5, 128, 32, 32, True, True), (4, 300, 16, 16, False, True)])
def test_one_launch_dl_step_equals_three_launch_step(built_lib, monkeypatch, S, B, H, W, interleaved, u8):
    """vaemdl_dlogistic_iwae_step: the cooperative one-launch kernel (forward with the unscaled derivatives parked in
    shared memory -> grid barrier -> finish -> grid barrier -> scale and store) agrees with forward + finish + backward
    as three launches to float32 round-off, and matches the float64 oracle (models/model03.py shapes, both layouts)."""
    from vae_mdl_b200 import functional as F
    g = torch.Generator().manual_seed(4100 + S + B + H)
    both = torch.randn(S, B, H, W, 6, generator=g)
    both[..., :3] = torch.rand(S, B, H, W, 3, generator=g)
    x_u8 = torch.randint(0, 256, (B, H, W, 3), dtype=torch.uint8, generator=g)
    x_u8.view(-1)[::13] = 0
    x_u8.view(-1)[3::19] = 255
    loc64 = both[..., :3].double().requires_grad_(True)
    ls64 = both[..., 3:].double().requires_grad_(True)
    x64 = O.normalize_u8(x_u8, torch.float64)
    ll64 = O.dlogistic_log_prob(x64, loc64, ls64, 0.0, 1.0, 256.0).sum((-1, -2, -3))
    extra = (ll64.detach().mean(0, keepdim=True) - ll64.detach()) + torch.randn(S, B, generator=g).double()
    loss64 = -O.logmeanexp(ll64 + extra.float().double(), 0).mean()
    loss64.backward()
    bd = both.to(DEV)
    if interleaved:
        loc, ls = bd[..., :3], bd[..., 3:]
    else:
        loc, ls = bd[..., :3].contiguous(), bd[..., 3:].contiguous()
    xd = x_u8.to(DEV) if u8 else (x_u8.float() / 255.0).to(DEV)
    ed = extra.float().to(DEV)
    monkeypatch.setenv("VAEMDL_FUSED", "1")
    a = F.dlogistic_iwae_step(loc, ls, xd, ed, 0.0, 1.0, 256.0)
    torch.cuda.synchronize()
    assert a[-1] == 1, "the one-launch kernel was not taken"
    monkeypatch.setenv("VAEMDL_FUSED", "0")
    b = F.dlogistic_iwae_step(loc, ls, xd, ed, 0.0, 1.0, 256.0)
    torch.cuda.synchronize()
    assert b[-1] == 3
    # the one-launch kernel takes the forward value out of the GRADIENT instantiation of the element function (evaluated
    # once): the compiler contracts a few FMAs differently there, so the two routes agree to float32 round-off, not bit
    # for bit; the importance weights see the ~1e-4 nat differences of the per-image sums
    for name, u, v in zip(("ll64", "log_w", "lme_b", "elbo", "g_ll", "dloc", "dls"), a[:-1], b[:-1]):
        u, v = u.double(), v.double()
        tol = {"ll64": 1e-7, "log_w": 1e-6, "lme_b": 1e-6, "elbo": 1e-6, "g_ll": 2e-3, "dloc": 2e-3, "dls": 2e-3}[name]
        assert ((u - v).norm() / v.norm()).item() <= tol, f"{name} differs between the one-launch and the three-launch step"
    assert ((a[0].cpu() - ll64.detach()).abs() / ll64.detach().abs()).max().item() <= 1e-5
    assert abs(-a[3].item() - loss64.item()) <= 1e-5 * abs(loss64.item())
    assert ((a[5].cpu().double() - loc64.grad).norm() / loc64.grad.norm()).item() <= 1e-4
    assert ((a[6].cpu().double() - ls64.grad).norm() / ls64.grad.norm()).item() <= 1e-4
