"""GPU parity tests of the MoDL kernels (through the C ABI) against the float64 oracle and the golden fixtures."""
import ctypes
import math

import numpy as np
import pytest
import torch

import oracle as O
from util import (GRAD_RTOL, LL_RTOL, assert_grad_close, assert_ll_close, canonical, golden, oracle_ll_and_grad, relnorm,
                  threshold_ambiguous, trained_like)

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


@pytest.fixture(scope="module")
def V(built_lib):
    import vae_mdl_b200
    return vae_mdl_b200


@pytest.fixture(scope="module")
def F(built_lib):
    from vae_mdl_b200 import functional
    return functional


# ------------------------------------------------------------------------------------------------
# golden fixtures
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,M", [("modl_m10_randn", 10), ("modl_m5_trained", 5), ("modl_m30_randn", 30), ("modl_m7_ragged", 7)])
def test_golden_forward_backward(F, V, name, M):
    z = golden(name)
    params = torch.from_numpy(z["params"]).to(DEV)
    x_u8 = torch.from_numpy(z["x_u8"]).to(DEV)
    ll = F.modl_log_likelihood(params, x_u8)
    assert_ll_close(ll, z["ll"])
    ll64 = F.modl_log_likelihood(params, x_u8, dtype=torch.float64)
    assert_ll_close(ll64, z["ll"], rtol=2e-7)
    lp = F.modl_log_prob(params, x_u8)
    assert (lp.cpu().double() - torch.from_numpy(z["lp"])).abs().max().item() < 5e-5
    # gradient kernel with the fixture's upstream weights
    dp = F.modl_backward(params, x_u8, g_image=torch.from_numpy(z["g_image"]).to(DEV))
    assert_grad_close(dp, z["grad_fixed"], M)
    # whole IWAE chain (comparable importance weights)
    loss, lpxz, dp2 = V.modl_iwae_step(params, x_u8, torch.from_numpy(z["extra"]).to(DEV))
    assert abs(loss.item() - float(z["loss"])) <= LL_RTOL * abs(float(z["loss"]))
    assert_grad_close(dp2, z["grad_iwae"], M)


# ------------------------------------------------------------------------------------------------
# seeded comparisons: every tiled instantiation, the any-M kernel, ragged shapes
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("S,B,H,W,M", [
    (2, 3, 8, 8, 5), (2, 3, 8, 8, 10), (2, 2, 8, 8, 20), (2, 2, 8, 8, 30),      # tiled kernels
    (1, 2, 4, 4, 1), (2, 2, 4, 4, 3), (1, 2, 4, 4, 64),                           # any-M kernel
    (3, 1, 5, 7, 5), (1, 3, 3, 3, 10), (2, 1, 7, 3, 30), (1, 1, 1, 1, 10),      # ragged: H*W % 32 != 0, odd pixel counts
    (2, 2, 32, 32, 10), (1, 2, 16, 16, 30),
])
def test_forward_backward_vs_oracle(F, S, B, H, W, M):
    params, x_u8, g = canonical(1000 + 37 * M + H, S, B, H, W, M)
    g_image = torch.randn(S, B, generator=g)
    lp64, ll64, grad64 = oracle_ll_and_grad(params, x_u8, g_image)
    pd, xd = params.to(DEV), x_u8.to(DEV)
    assert_ll_close(F.modl_log_likelihood(pd, xd), ll64)
    assert_ll_close(F.modl_log_likelihood(pd, xd, dtype=torch.float64), ll64, rtol=5e-7)
    lp = F.modl_log_prob(pd, xd)
    assert lp.shape == (S, B, H, W)
    assert (lp.cpu().double() - lp64).abs().max().item() < 5e-5
    assert_grad_close(F.modl_backward(pd, xd, g_image=g_image.to(DEV)), grad64, M)


@pytest.mark.parametrize("M", [5, 10, 30])
def test_trained_like_distribution(F, M):
    """Narrow scales (raw log-scale ~ N(-3,1)): edge and low-probability branches fire constantly and exp(-h) leaves
    the polynomial range, so the MUFU path of every sub-pixel term is exercised."""
    S, B, H, W = 2, 3, 16, 16
    params, x_u8, g = trained_like(77 + M, S, B, H, W, M)
    g_image = torch.randn(S, B, generator=g)
    lp64, ll64, grad64 = oracle_ll_and_grad(params, x_u8, g_image)
    amb = threshold_ambiguous(params.double(), O.normalize_u8(x_u8, torch.float64))
    pd, xd = params.to(DEV), x_u8.to(DEV)
    ll = F.modl_log_likelihood(pd, xd, dtype=torch.float64).cpu()
    ok = ~amb
    assert ok.any()
    assert (((ll - ll64).abs() / ll64.abs())[ok]).max().item() <= LL_RTOL
    if not amb.any():
        assert_grad_close(F.modl_backward(pd, xd, g_image=g_image.to(DEV)), grad64, M)


def test_x_dtypes_and_ranges_agree(F):
    """uint8 bytes, float32 in [0,1] and float32 in [-1,1] (utils/mdl_openai.py) are the same data."""
    from vae_mdl_b200 import _abi
    params, x_u8, _ = canonical(5, 2, 3, 8, 8, 10)
    pd = params.to(DEV)
    x01 = O.normalize_u8(x_u8)
    a = F.modl_log_likelihood(pd, x_u8.to(DEV))
    b = F.modl_log_likelihood(pd, x01.to(DEV))
    c = F.modl_log_likelihood(pd, (x01 * 2 - 1).to(DEV), _abi.RANGE_SYM, _abi.EDGE_OPENAI)
    assert torch.equal(a, b) and torch.equal(a, c)


def test_edge_modes_differ_only_on_unbinned_data(F):
    """x in (-1, -0.999): utils/mdl.py:200 uses the normal branch, utils/mdl_openai.py:139 the left-edge branch."""
    from vae_mdl_b200 import _abi
    params, _, g = canonical(6, 1, 2, 4, 4, 5)
    x = torch.rand(2, 4, 4, 3, generator=g) * 2 - 1
    x[0, 0, 0, 0] = -0.9995
    x[1, 1, 1, 2] = 0.9995
    pd, xd = params.to(DEV), x.to(DEV)
    lp_mdl = F.modl_log_prob(pd, xd, _abi.RANGE_SYM, _abi.EDGE_MDL).cpu().double()
    lp_oai = F.modl_log_prob(pd, xd, _abi.RANGE_SYM, _abi.EDGE_OPENAI).cpu().double()
    want_oai = O.modl_openai_log_prob(params[0].double(), x.double())
    assert (lp_oai[0] - want_oai).abs().max().item() < 5e-5
    # mdl.py semantics on x already in [-1,1]: feed (x+1)/2 to the oracle, whose 2x-1 is then exact enough in float64
    want_mdl = O.modl_log_prob(params.double(), (x.double() + 1) / 2)[0, ..., 0]
    assert (lp_mdl[0] - want_mdl).abs().max().item() < 5e-5
    diff = (lp_mdl - lp_oai).abs()[0]
    assert diff[0, 0, 0].item() > 1e-3 and diff[1, 1, 1].item() > 1e-3
    diff[0, 0, 0] = 0
    diff[1, 1, 1] = 0
    assert diff.max().item() == 0.0


def test_broadcast_x_without_batch_dim(F):
    """models/model05.py:173: x [H,W,3] against parameters [S,1,H,W,10M]."""
    params, x_u8, _ = canonical(8, 6, 1, 8, 8, 5)
    a = F.modl_log_likelihood(params.to(DEV), x_u8[0].to(DEV))
    b = F.modl_log_likelihood(params.to(DEV), x_u8.to(DEV))
    assert a.shape == (6, 1) and torch.equal(a, b)
    with pytest.raises(ValueError):
        F.modl_log_likelihood(params.to(DEV), torch.zeros(4, 8, 8, 3, device=DEV))  # 6 parameter images vs 4 observed


def test_log_domain_fallback_extreme_tail(F):
    """Every mixture is ~80 scale units away from x: the linear-domain mixture sum underflows float32 and the kernel
    must redo the pixel in the log domain (value ~ -250 nats per pixel)."""
    M = 10
    S, B, H, W = 1, 2, 8, 8
    g = torch.Generator().manual_seed(3)
    params = torch.randn(S, B, H, W, 10 * M, generator=g) * 0.05
    for c in range(3):
        params[..., M + c * 3 * M: M + c * 3 * M + M] += 0.95            # loc ~ +0.95
        params[..., M + c * 3 * M + M: M + c * 3 * M + 2 * M] -= 4.6     # narrow scales (exp(4.6) ~ 100)
    x_u8 = torch.randint(0, 40, (B, H, W, 3), dtype=torch.uint8, generator=g)  # dark pixels, far from loc
    g_image = torch.randn(S, B, generator=g)
    lp64, ll64, grad64 = oracle_ll_and_grad(params, x_u8, g_image)
    assert lp64.min().item() < -200
    pd, xd = params.to(DEV), x_u8.to(DEV)
    lp = F.modl_log_prob(pd, xd).cpu().double()
    assert torch.isfinite(lp).all()
    assert ((lp - lp64).abs() / lp64.abs()).max().item() < 1e-5
    assert_ll_close(F.modl_log_likelihood(pd, xd), ll64)
    assert_grad_close(F.modl_backward(pd, xd, g_image=g_image.to(DEV)), grad64, M)


def test_logscale_clamp_masks_gradient(F):
    M = 10
    params, x_u8, g = canonical(12, 1, 2, 4, 4, M)
    params[..., 2 * M + 3] = -9.0                 # sR of mixture 3 below the clamp everywhere
    params[..., 8 * M + 6] = -7.5                 # sB of mixture 6 below the clamp everywhere
    params[0, 0, 0, 0, 5 * M: 6 * M] = -7.0       # sG exactly on the clamp: gradient flows (tf.maximum tie rule)
    g_image = torch.ones(1, 2)
    _, ll64, grad64 = oracle_ll_and_grad(params, x_u8, g_image)
    dp = F.modl_backward(params.to(DEV), x_u8.to(DEV), g_image=g_image.to(DEV)).cpu()
    assert (dp[..., 2 * M + 3] == 0).all() and (dp[..., 8 * M + 6] == 0).all()
    assert (grad64[0, 0, 0, 0, 5 * M: 6 * M] != 0).any()
    assert (dp[0, 0, 0, 0, 5 * M: 6 * M] != 0).any()
    assert_grad_close(dp, grad64, M)


def test_per_pixel_upstream_gradient_and_autograd(V):
    """log_prob(x) [...,H,W,1] with an arbitrary upstream gradient (the generic drop-in path through autograd)."""
    S, B, H, W, M = 2, 2, 8, 8, 10
    params, x_u8, g = canonical(13, S, B, H, W, M)
    w = torch.randn(S, B, H, W, 1, generator=g)
    p64 = params.double().requires_grad_(True)
    (O.modl_log_prob(p64, O.normalize_u8(x_u8, torch.float64)) * w.double()).sum().backward()
    pd = params.to(DEV).requires_grad_(True)
    dist = V.MixtureDiscretizedLogistic(pd)
    out = dist.log_prob(O.normalize_u8(x_u8).to(DEV))
    assert out.shape == (S, B, H, W, 1)
    (out * w.to(DEV)).sum().backward()
    assert_grad_close(pd.grad, p64.grad, M)


# ------------------------------------------------------------------------------------------------
# full-size (BASELINE config 1) size-independent properties
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("M,B", [(10, 64), (5, 128)])
def test_config1_full_size_properties(F, M, B):
    S, H, W = 5, 32, 32
    gen = torch.Generator(device=DEV).manual_seed(1234)
    params = torch.randn(S, B, H, W, 10 * M, device=DEV, generator=gen)
    x_u8 = torch.randint(0, 256, (B, H, W, 3), dtype=torch.uint8, device=DEV, generator=gen)
    g_image = torch.randn(S, B, device=DEV, generator=gen)
    ll = F.modl_log_likelihood(params, x_u8)
    lp = F.modl_log_prob(params, x_u8)
    # (a) the fused per-image sum equals the sum of the per-pixel output
    assert ((lp.double().sum((-1, -2)) - ll.double()).abs() / ll.double().abs()).max().item() < 1e-6
    # (b) bitwise reproducible
    assert torch.equal(ll, F.modl_log_likelihood(params, x_u8))
    dp = F.modl_backward(params, x_u8, g_image=g_image)
    assert torch.equal(dp, F.modl_backward(params, x_u8, g_image=g_image))
    # (c) the gradient is linear in the upstream weights: exact for a power-of-two factor, up to values that
    #     flush to zero (the kernels run with denormals flushed)
    assert (F.modl_backward(params, x_u8, g_image=2 * g_image) - 2 * dp).abs().max().item() < 1e-36
    # (d) mixture-logit gradients of a pixel sum to zero (responsibilities and softmax both sum to one)
    s = dp[..., :M].sum(-1)
    assert s.abs().max().item() < 1e-5 * g_image.abs().max().item()
    # (e) utils/mdl.py semantics == utils/mdl_openai.py semantics on binned data
    from vae_mdl_b200 import _abi
    assert torch.equal(ll, F.modl_log_likelihood(params, x_u8, _abi.RANGE_UNIT, _abi.EDGE_OPENAI))
    # (f) spot-check 4 images against the oracle
    idx = [(0, 0), (1, 7), (4, B - 1), (2, B // 2)]
    for s_i, b_i in idx:
        want = O.modl_log_prob(params[s_i, b_i].cpu().double()[None], O.normalize_u8(x_u8[b_i].cpu(), torch.float64)[None]).sum()
        assert abs(ll[s_i, b_i].item() - want.item()) <= LL_RTOL * abs(want.item())


def test_abi_direct_call_and_errors(built_lib):
    """The raw C entry point, without the Python wrappers."""
    L = built_lib
    S, B, H, W, M = 2, 2, 8, 8, 10
    params, x_u8, _ = canonical(21, S, B, H, W, M)
    pd, xd = params.to(DEV), x_u8.to(DEV)
    ll = torch.empty(S, B, device=DEV)
    nb = L.vaemdl_modl_workspace_bytes(S * B, H, W)
    ws = torch.empty(nb // 8 + 1, dtype=torch.float64, device=DEV)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    rc = L.vaemdl_modl_fwd(pd.data_ptr(), xd.data_ptr(), 1, 0, 0, S * B, B, H, W, M, None, ll.data_ptr(), None,
                           ws.data_ptr(), nb, st)
    assert rc == 0
    torch.cuda.synchronize()
    want = O.modl_log_prob(params.double(), O.normalize_u8(x_u8, torch.float64)).sum((-1, -2, -3))
    assert_ll_close(ll, want)
    # workspace too small
    assert L.vaemdl_modl_fwd(pd.data_ptr(), xd.data_ptr(), 1, 0, 0, S * B, B, H, W, M, None, ll.data_ptr(), None,
                             ws.data_ptr(), 8, st) == -4
    # misaligned parameter pointer
    assert L.vaemdl_modl_fwd(pd.data_ptr() + 4, xd.data_ptr(), 1, 0, 0, S * B, B, H, W, M, None, ll.data_ptr(), None,
                             ws.data_ptr(), nb, st) == -2


@pytest.mark.parametrize("force", ["0", "1"])
@pytest.mark.parametrize("name", ["modl_m5_trained"])
def test_both_n_mix_5_kernels_on_the_golden_fixture(F, V, monkeypatch, force, name):
    """n_mix = 5 has two instantiations (component pairs / pixel pairs, chosen by problem size): both must reproduce the
    fixture, including its pixels whose mixture sum underflows float32 (log-domain fallback in one half of a lane)."""
    monkeypatch.setenv("VAEMDL_PP", force)
    z = golden(name)
    params = torch.from_numpy(z["params"]).to(DEV)
    x_u8 = torch.from_numpy(z["x_u8"]).to(DEV)
    assert_ll_close(F.modl_log_likelihood(params, x_u8, dtype=torch.float64), z["ll"], rtol=2e-7)
    lp = F.modl_log_prob(params, x_u8)
    assert (lp.cpu().double() - torch.from_numpy(z["lp"])).abs().max().item() < 5e-5
    dp = F.modl_backward(params, x_u8, g_image=torch.from_numpy(z["g_image"]).to(DEV))
    assert_grad_close(dp, z["grad_fixed"], 5)
    loss, lpxz, dp2 = V.modl_iwae_step(params, x_u8, torch.from_numpy(z["extra"]).to(DEV))
    assert abs(loss.item() - float(z["loss"])) <= LL_RTOL * abs(float(z["loss"]))
    assert_grad_close(dp2, z["grad_iwae"], 5)


@pytest.mark.parametrize("S,B,H,W,M,force", [(2, 3, 8, 8, 5, "1"), (3, 1, 5, 7, 5, "1"), (1, 2, 9, 9, 5, "1"), (2, 2, 16, 12, 5, "1"),
                                               (1, 2, 9, 9, 2, None), (2, 1, 10, 7, 4, None), (1, 3, 8, 8, 6, None),
                                               (2, 2, 8, 9, 8, None), (1, 1, 12, 12, 9, None)])
def test_pixel_pair_kernel_vs_oracle(F, monkeypatch, S, B, H, W, M, force):
    """The pixel-pair kernel (n_mix 1..9): ragged tiles, tiles straddling images, images smaller than a tile, with the
    trained-like distribution (narrow scales, low-probability branch, underflowing mixture sums)."""
    if force is not None:
        monkeypatch.setenv("VAEMDL_PP", force)
    params, x_u8, g = trained_like(500 + 11 * M + H, S, B, H, W, M)
    x_u8[0, 0, 0] = torch.tensor([0, 255, 0], dtype=torch.uint8)      # both edges in one pixel
    g_image = torch.randn(S, B, generator=g)
    lp64, ll64, grad64 = oracle_ll_and_grad(params, x_u8, g_image)
    ok = ~threshold_ambiguous(params.double(), O.normalize_u8(x_u8, torch.float64))
    pd, xd = params.to(DEV), x_u8.to(DEV)
    ll = F.modl_log_likelihood(pd, xd, dtype=torch.float64).cpu()
    assert ((ll - ll64).abs() / ll64.abs())[ok].max().item() <= LL_RTOL
    dp = F.modl_backward(pd, xd, g_image=g_image.to(DEV)).cpu().double()
    if bool(ok.all()):
        assert_grad_close(dp, grad64, M)
    else:
        assert relnorm(dp[ok], grad64[ok]) <= GRAD_RTOL


@pytest.mark.parametrize("S,B,H,W,M", [(2, 2, 8, 8, 11), (1, 3, 5, 7, 12), (2, 1, 9, 9, 13), (1, 2, 8, 8, 16), (2, 2, 6, 7, 18),
                                        (1, 2, 8, 8, 25), (1, 2, 5, 5, 32), (2, 1, 8, 8, 40), (1, 2, 4, 6, 50), (1, 1, 8, 8, 64),
                                        (1, 2, 2, 2, 15), (1, 1, 1, 1, 24), (3, 2, 16, 12, 14), (1, 3, 3, 3, 21)])
@pytest.mark.parametrize("plain", [False, True])
def test_runtime_tiled_kernel_vs_oracle(F, S, B, H, W, M, plain):
    """n_mix without its own instantiation runs on modl_rt_kernel (run-time split of a pixel over LPP lanes, uneven and
    empty component chunks, padding pairs, rotated component order): ragged tiles, tiles straddling images, images
    smaller than a tile (atomics route), trained-like distribution; both mean chains."""
    params, x_u8, g = trained_like(900 + 13 * M + H, S, B, H, W, M)
    x_u8[0, 0, 0] = torch.tensor([0, 255, 0], dtype=torch.uint8)
    g_image = torch.randn(S, B, generator=g)
    pd, xd = params.to(DEV), x_u8.to(DEV)
    if plain:
        p64 = params.double().requires_grad_(True)
        lp = O.mdl_plain_log_prob(p64, O.normalize_u8(x_u8, torch.float64))
        ll64 = lp.sum((-1, -2))
        (ll64 * g_image.double()).sum().backward()
        ll64, grad64 = ll64.detach(), p64.grad
        ll = F.modl_log_likelihood(pd, xd, dtype=torch.float64, plain=True).cpu()
        assert ((ll - ll64).abs() / ll64.abs()).max().item() <= LL_RTOL
        dp = F.modl_backward(pd, xd, g_image=g_image.to(DEV), plain=True).cpu().double()
        assert relnorm(dp, grad64) <= GRAD_RTOL
        return
    lp64, ll64, grad64 = oracle_ll_and_grad(params, x_u8, g_image)
    ok = ~threshold_ambiguous(params.double(), O.normalize_u8(x_u8, torch.float64))
    assert bool(ok.any()), "pick another seed: every image holds a threshold-ambiguous sub-pixel"
    ll = F.modl_log_likelihood(pd, xd, dtype=torch.float64).cpu()
    assert ((ll - ll64).abs() / ll64.abs())[ok].max().item() <= LL_RTOL
    lp = F.modl_log_prob(pd, xd).cpu().double()
    assert ((lp - lp64).abs().flatten(2).amax(-1))[ok].max().item() < 5e-5
    dp = F.modl_backward(pd, xd, g_image=g_image.to(DEV)).cpu().double()
    if bool(ok.all()):
        assert_grad_close(dp, grad64, M)
    else:
        assert relnorm(dp[ok], grad64[ok]) <= GRAD_RTOL


@pytest.mark.parametrize("S,B,H,W,M", [(2, 3, 8, 8, 10), (3, 1, 5, 7, 5), (1, 3, 3, 3, 10), (2, 1, 7, 3, 30), (1, 2, 9, 9, 2),
                                        (2, 2, 5, 5, 7), (1, 1, 1, 1, 10), (5, 2, 16, 12, 5), (2, 2, 6, 6, 20), (1, 2, 4, 4, 13),
                                        (2, 2, 7, 5, 11), (1, 3, 6, 6, 16), (2, 1, 5, 9, 25), (1, 2, 3, 5, 64), (1, 1, 9, 9, 8)])
def test_no_write_outside_the_callers_buffers(built_lib, S, B, H, W, M):
    """Guard words around every output and around the workspace (exactly vaemdl_modl_workspace_bytes long) must survive
    the forward, fused-finish and backward launches on ragged shapes (compute-sanitizer is not available on this pool)."""
    L = built_lib
    n_img, n_px = S * B, S * B * H * W
    g = torch.Generator().manual_seed(S + B + H + W + M)
    params = torch.randn(S, B, H, W, 10 * M, generator=g).to(DEV)
    x = torch.randint(0, 256, (B, H, W, 3), dtype=torch.uint8, generator=g).to(DEV)
    G = 64  # guard elements on each side

    def guarded(n, dtype, fill):
        buf = torch.full((n + 2 * G,), fill, dtype=dtype, device=DEV)
        return buf, buf[G:G + n]

    ws_bytes = L.vaemdl_modl_workspace_bytes(n_img, H, W)
    ws_buf, ws = guarded((ws_bytes + 7) // 8, torch.float64, -7.0)
    lp_buf, lp = guarded(n_px, torch.float32, -7.0)
    ll_buf, ll = guarded(n_img, torch.float32, -7.0)
    ll64_buf, ll64 = guarded(n_img, torch.float64, -7.0)
    lw_buf, lw = guarded(n_img, torch.float32, -7.0)
    gl_buf, gl = guarded(n_img, torch.float32, -7.0)
    lme_buf, lme = guarded(B, torch.float32, -7.0)
    el_buf, el = guarded(1, torch.float32, -7.0)
    dp_buf, dp = guarded(n_px * 10 * M + 0, torch.float32, -7.0)
    assert ws.data_ptr() % 8 == 0
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    if dp.data_ptr() % 16:  # the gradient pointer must be 16-byte aligned: shift the view inside the guard band
        pytest.skip("allocator returned an unaligned guard view")
    rc = L.vaemdl_modl_fwd(params.data_ptr(), x.data_ptr(), 1, 0, 0, n_img, B, H, W, M, lp.data_ptr(), ll.data_ptr(),
                           ll64.data_ptr(), ws.data_ptr(), ws_bytes, st)
    assert rc == 0
    rc = L.vaemdl_modl_iwae_fwd(params.data_ptr(), x.data_ptr(), 1, 0, 0, S, B, 0, B, H, W, M, None, ll.data_ptr(),
                                ll64.data_ptr(), lw.data_ptr(), lme.data_ptr(), el.data_ptr(), gl.data_ptr(), ws.data_ptr(),
                                ws_bytes, st)
    assert rc == 0
    rc = L.vaemdl_modl_bwd(params.data_ptr(), x.data_ptr(), 1, 0, 0, n_img, B, H, W, M, gl.data_ptr(), None, dp.data_ptr(), st)
    assert rc == 0
    torch.cuda.synchronize()
    for name, buf in [("ws", ws_buf), ("lp", lp_buf), ("ll", ll_buf), ("ll64", ll64_buf), ("log_w", lw_buf), ("g_ll", gl_buf),
                      ("lme", lme_buf), ("elbo", el_buf), ("dparams", dp_buf)]:
        assert bool((buf[:G] == -7.0).all()) and bool((buf[-G:] == -7.0).all()), f"guard band of {name} was overwritten"
    assert not bool(torch.isnan(dp).any()) and bool((lp != -7.0).all()) and bool((dp != -7.0).any())


def test_every_byte_value_decodes_exactly(F, V):
    """uint8 x and the float image x/255 (torch division, correctly rounded) give bit-identical results for all 256
    byte values in every channel -- MoDL and plain DL, element-wise outputs."""
    M = 10
    g = torch.Generator().manual_seed(256)
    k = torch.arange(256, dtype=torch.uint8)
    x_u8 = torch.stack([k, k.flip(0), (k.int() * 7 % 256).to(torch.uint8)], -1).reshape(1, 16, 16, 3)
    params = torch.randn(2, 1, 16, 16, 10 * M, generator=g).to(DEV)
    xd = x_u8.to(DEV)
    x01 = (x_u8.float() / 255.0).to(DEV)
    assert torch.equal(F.modl_log_prob(params, xd), F.modl_log_prob(params, x01))
    p5 = torch.randn(2, 1, 16, 16, 50, generator=g).to(DEV)
    assert torch.equal(F.modl_log_prob(p5, xd), F.modl_log_prob(p5, x01))
    both = torch.randn(2, 1, 16, 16, 6, generator=g).to(DEV)
    mu, ls = torch.split(both, 3, dim=-1)
    d = V.DiscretizedLogistic(mu, ls, low=0.0, high=1.0, levels=256.0)
    assert torch.equal(d.log_prob(xd), d.log_prob(x01))
