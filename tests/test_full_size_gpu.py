"""BASELINE.json's full sizes on the GPU: size-independent properties + oracle spot checks (the oracle cannot score these
shapes whole in seconds).  configs[1] (plain DL), configs[2] (sampler), configs[3] (5000-sample evaluation, one image) and
the per-GPU shard of configs[4] (64x64x3, 10 mixtures, 16 x 32); configs[0] is in test_modl_gpu.py."""
import math

import pytest
import torch

import oracle as O
from util import GRAD_RTOL, LL_RTOL, relnorm

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


def test_config5_shard_properties(F, V):
    S, B, H, W, M = 16, 32, 64, 64, 10
    gen = torch.Generator(device=DEV).manual_seed(5)
    params = torch.randn(S, B, H, W, 10 * M, device=DEV, generator=gen)
    x_u8 = torch.randint(0, 256, (B, H, W, 3), dtype=torch.uint8, device=DEV, generator=gen)
    ll64 = F.modl_log_likelihood(params, x_u8, dtype=torch.float64)
    lp = F.modl_log_prob(params, x_u8)
    # (a) fused per-image sums == sums of the per-pixel output (float64 accumulation of the same float32 values)
    assert ((lp.double().sum((-1, -2)) - ll64).abs() / ll64.abs()).max().item() < 1e-12
    # (b) the fused finish gives the same sums bit for bit, and the IWAE tail of models/loss.py:34-37 on top of them
    extra = (ll64.mean(0, keepdim=True) - ll64).float() + torch.randn(S, B, device=DEV, generator=gen)
    ll_f, log_w, lme_b, elbo, g_ll = F.modl_iwae_forward(params, x_u8, extra)
    assert torch.equal(ll_f, ll64)
    lw = ll64 + extra.double()
    want_lme = torch.logsumexp(lw, 0) - math.log(S)
    assert ((lme_b.double() - want_lme).abs() / want_lme.abs()).max().item() < 1e-6
    assert abs(elbo.item() - want_lme.mean().item()) <= 1e-6 * abs(want_lme.mean().item())
    assert relnorm(g_ll, -torch.softmax(lw, 0) / B) < 1e-5
    assert abs(g_ll.sum().item() + 1.0) < 1e-5                      # the softmax weights of every image sum to one
    # (c) splitting the batch changes nothing (images are independent): the shard of ranks 0/1 of a 2-GPU run
    h = B // 2
    l0, _, _, e0, g0 = F.modl_iwae_forward(params[:, :h].contiguous(), x_u8[:h], extra[:, :h].contiguous(), b_total=B)
    l1, _, _, e1, g1 = F.modl_iwae_forward(params[:, h:].contiguous(), x_u8[h:], extra[:, h:].contiguous(), b_total=B)
    assert ((torch.cat([l0, l1], 1) - ll64).abs() / ll64.abs()).max().item() < 1e-12
    assert abs((e0 + e1).item() - elbo.item()) <= 1e-6 * abs(elbo.item())
    assert relnorm(torch.cat([g0, g1], 1), g_ll) < 1e-6
    # (d) gradient: reproducible, linear in the upstream weights, logit gradients of a pixel sum to zero
    dp = F.modl_backward(params, x_u8, g_image=g_ll)
    assert torch.equal(dp, F.modl_backward(params, x_u8, g_image=g_ll))
    assert (F.modl_backward(params, x_u8, g_image=4 * g_ll) - 4 * dp).abs().max().item() < 1e-36
    assert dp[..., :M].sum(-1).abs().max().item() < 1e-5 * g_ll.abs().max().item()
    loss, lpxz, dp2 = V.modl_iwae_step(params, x_u8, extra)
    assert torch.equal(dp2, dp) and torch.equal(lpxz, ll64)
    # (e) oracle spot checks: two whole images, log-likelihood and gradient
    for s_i, b_i in [(0, 0), (S - 1, B - 1)]:
        p64 = params[s_i, b_i].cpu().double()[None].requires_grad_(True)
        x64 = O.normalize_u8(x_u8[b_i].cpu(), torch.float64)[None]
        want = O.modl_log_prob(p64, x64).sum()
        assert abs(ll64[s_i, b_i].item() - want.item()) <= LL_RTOL * abs(want.item())
        (want * g_ll[s_i, b_i].item()).backward()
        assert relnorm(dp[s_i, b_i], p64.grad[0]) <= GRAD_RTOL


def test_config4_one_image_5000_importance_samples(F):
    """models/model05.py:168-176 for one test image: [5000, 1, 32, 32, 100] parameters, x broadcast over the samples."""
    S, H, W, M = 5000, 32, 32, 10
    gen = torch.Generator(device=DEV).manual_seed(4)
    params = torch.randn(S, 1, H, W, 10 * M, device=DEV, generator=gen)
    x_u8 = torch.randint(0, 256, (1, H, W, 3), dtype=torch.uint8, device=DEV, generator=gen)
    ll_f, log_w, lme_b, elbo, g_ll = F.modl_iwae_forward(params, x_u8, None)
    want = torch.logsumexp(ll_f[:, 0], 0) - math.log(S)
    assert abs(lme_b[0].item() - want.item()) <= 1e-6 * abs(want.item()) and abs(elbo.item() - want.item()) <= 1e-6 * abs(want.item())
    # streaming the samples in chunks of 250 (tools / dist.IwaeEvaluator) gives the same per-sample sums
    chunks = [F.modl_log_likelihood(params[s0:s0 + 250], x_u8, dtype=torch.float64) for s0 in range(0, S, 250)]
    assert ((torch.cat(chunks, 0) - ll_f).abs() / ll_f.abs()).max().item() < 1e-12
    # x without its batch dim (models/model05.py:173 passes [H,W,3]) is the same thing
    assert torch.equal(F.modl_log_likelihood(params, x_u8[0], dtype=torch.float64), ll_f)
    # oracle spot check: 8 of the 5000 samples
    idx = torch.tensor([0, 1, 777, 2499, 2500, 4096, 4998, 4999])
    p64 = params[idx, 0].cpu().double()
    want8 = O.modl_log_prob(p64, O.normalize_u8(x_u8[0].cpu(), torch.float64)).sum((-1, -2, -3))
    assert ((ll_f[idx, 0].cpu() - want8).abs() / want8.abs()).max().item() <= LL_RTOL
    bpd = -lme_b[0].item() / (math.log(2.0) * H * W * 3)
    assert 8.0 < bpd < 11.0                                        # random parameters: ~9.5 bits/dim


def test_config2_plain_dl_full_size(F, V):
    S, B, H, W = 5, 128, 32, 32
    gen = torch.Generator(device=DEV).manual_seed(2)
    both = torch.randn(S, B, H, W, 6, device=DEV, generator=gen)
    both[..., :3].uniform_(generator=gen)
    mu, lstd = torch.split(both, 3, dim=-1)
    x_u8 = torch.randint(0, 256, (B, H, W, 3), dtype=torch.uint8, device=DEV, generator=gen)
    d = V.DiscretizedLogistic(mu, lstd, low=0.0, high=1.0, levels=256.0)
    ll64 = d.log_likelihood(x_u8, dtype=torch.float64)
    lp = d.log_prob(x_u8)
    # (a lane adds its six element values in float32 before the float64 accumulation: agreement to float32 rounding)
    assert ((lp.double().sum((-1, -2, -3)) - ll64).abs() / ll64.abs()).max().item() < 1e-7
    loss, lpxz, dloc, dls = V.dlogistic_iwae_step(mu, lstd, x_u8, None, 0.0, 1.0, 256.0)
    # (the one-launch step takes the forward value out of the gradient instantiation of the element function: float32
    # round-off apart from the forward kernel's, see tests/test_dlogistic_gpu.py)
    assert ((lpxz - ll64).abs() / ll64.abs()).max().item() < 1e-7
    ll64 = lpxz
    want = -(torch.logsumexp(ll64, 0) - math.log(S)).mean()
    assert abs(loss.item() - want.item()) <= 1e-6 * abs(want.item())
    loss2, _, dloc2, dls2 = V.dlogistic_iwae_step(mu, lstd, x_u8, None, 0.0, 1.0, 256.0)
    assert torch.equal(dloc, dloc2) and torch.equal(dls, dls2)
    # oracle spot checks on whole images: log-likelihood AND gradient.  With random parameters one importance sample
    # carries all the weight (the others' gradients underflow), so the gradient is checked on a second step whose `extra`
    # makes the weights of a batch element comparable -- upstream = that step's softmax weights.
    extra = (ll64.mean(0, keepdim=True) - ll64).float() + torch.randn(S, B, device=DEV, generator=gen)
    _, lpxz_b, dloc, dls = V.dlogistic_iwae_step(mu, lstd, x_u8, extra, 0.0, 1.0, 256.0)
    assert torch.equal(lpxz_b, ll64)
    g_ll = -torch.softmax(ll64 + extra.double(), 0) / B
    assert g_ll.abs().min().item() > 1e-8
    both_grad = torch.cat([dloc, dls], dim=-1)
    for s_i, b_i in [(0, 0), (4, 127), (2, 63)]:
        b64 = both[s_i, b_i].cpu().double()[None].requires_grad_(True)
        x64 = O.normalize_u8(x_u8[b_i].cpu(), torch.float64)[None]
        w = O.dlogistic_log_prob(x64, b64[..., :3], b64[..., 3:], 0.0, 1.0, 256.0).sum()
        assert abs(ll64[s_i, b_i].item() - w.item()) <= LL_RTOL * abs(w.item())
        (w * g_ll[s_i, b_i].item()).backward()
        assert relnorm(both_grad[s_i, b_i], b64.grad[0]) <= GRAD_RTOL
        assert relnorm(dloc[s_i, b_i], b64.grad[0][..., :3]) <= GRAD_RTOL and relnorm(dls[s_i, b_i], b64.grad[0][..., 3:]) <= GRAD_RTOL


def test_config3_sampler_full_size(V):
    N, H, W, M = 10000, 32, 32, 10
    gen = torch.Generator(device=DEV).manual_seed(3)
    l = torch.randn(N, H, W, 10 * M, device=DEV, generator=gen)
    um = torch.rand(N, H, W, M, device=DEV, generator=gen) * (1 - 2e-5) + 1e-5
    ul = torch.rand(N, H, W, 3, device=DEV, generator=gen) * (1 - 2e-5) + 1e-5
    x, xq, idx = V.sample_from_discretized_mix_logistic(l, M, um, ul, return_index=True, return_quantised=True)
    assert x.shape == (N, H, W, 3) and x.min().item() >= -1.0 and x.max().item() <= 1.0 and idx.max().item() < M
    # the quantised bytes are the rounded float output except within float32 rounding of a bin boundary
    q_from_x = torch.round(255.0 * (x.double() * 0.5 + 0.5)).to(torch.uint8)
    assert (q_from_x != xq).float().mean().item() < 1e-5
    # same call, same bits
    x2, xq2, idx2 = V.sample_from_discretized_mix_logistic(l, M, um, ul, return_index=True, return_quantised=True)
    assert torch.equal(xq, xq2) and torch.equal(idx, idx2) and torch.equal(x, x2)
    # bit-exact against the float64 oracle on 24 of the 10,000 images
    pick = torch.linspace(0, N - 1, 24).long()
    x64, idx64 = O.sample_from_discretized_mix_logistic(l[pick].cpu(), M, um[pick].cpu(), ul[pick].cpu())
    assert int((idx[pick].cpu().long() != idx64).sum()) == 0
    assert int((xq[pick].cpu() != O.quantise(x64 * 0.5 + 0.5)).sum()) == 0
    assert (x[pick].cpu().double() - x64).abs().max().item() < 1e-6


@pytest.mark.parametrize("M,B", [(16, 20), (12, 27), (40, 8)])
def test_runtime_tiled_kernel_full_size_properties(F, M, B):
    """An n_mix without its own instantiation (modl_rt_kernel) at the config-5 image size: per-pixel / per-image
    consistency, reproducibility, linearity, zero-sum logit gradients, oracle spot checks."""
    S, H, W = 16, 64, 64
    gen = torch.Generator(device=DEV).manual_seed(50 + M)
    params = torch.randn(S, B, H, W, 10 * M, device=DEV, generator=gen)
    x_u8 = torch.randint(0, 256, (B, H, W, 3), dtype=torch.uint8, device=DEV, generator=gen)
    ll64 = F.modl_log_likelihood(params, x_u8, dtype=torch.float64)
    lp = F.modl_log_prob(params, x_u8)
    assert ((lp.double().sum((-1, -2)) - ll64).abs() / ll64.abs()).max().item() < 1e-12
    assert torch.equal(ll64, F.modl_log_likelihood(params, x_u8, dtype=torch.float64))
    g = torch.randn(S, B, device=DEV, generator=gen)
    dp = F.modl_backward(params, x_u8, g_image=g)
    assert torch.equal(dp, F.modl_backward(params, x_u8, g_image=g))
    assert (F.modl_backward(params, x_u8, g_image=4 * g) - 4 * dp).abs().max().item() < 1e-30
    assert dp[..., :M].sum(-1).abs().max().item() < 2e-5 * g.abs().max().item()
    for s_i, b_i in [(0, 0), (S - 1, B - 1), (7, B // 2)]:
        p64 = params[s_i, b_i].cpu().double()[None].requires_grad_(True)
        x64 = O.normalize_u8(x_u8[b_i].cpu(), torch.float64)[None]
        want = O.modl_log_prob(p64, x64).sum()
        assert abs(ll64[s_i, b_i].item() - want.item()) <= LL_RTOL * abs(want.item())
        (want * g[s_i, b_i].item()).backward()
        assert relnorm(dp[s_i, b_i], p64.grad[0]) <= GRAD_RTOL


@pytest.mark.parametrize("S,B,M", [(5, 64, 10), (5, 128, 5)])
def test_config1_one_launch_step_full_size(F, V, monkeypatch, S, B, M):
    """BASELINE configs[0] (and the reference's real default, n_mix 5 x batch 128) through vaemdl_modl_iwae_step: the
    cooperative one-launch kernel is taken, equals the three-launch route bit for bit, and matches the oracle on whole
    images picked from both ends of the tensor."""
    H = W = 32
    gen = torch.Generator(device=DEV).manual_seed(100 + M)
    params = torch.randn(S, B, H, W, 10 * M, device=DEV, generator=gen)
    x_u8 = torch.randint(0, 256, (B, H, W, 3), dtype=torch.uint8, device=DEV, generator=gen)
    extra = torch.randn(S, B, device=DEV, generator=gen) * 3
    monkeypatch.delenv("VAEMDL_FUSED", raising=False)
    default = F.modl_iwae_step(params, x_u8, extra)
    # three launches are the default (faster since the gradient kernel overlaps the finish kernel); the cooperative
    # one-launch kernel is opt-in (VAEMDL_FUSED=1)
    assert default[-1] == 3
    monkeypatch.setenv("VAEMDL_FUSED", "1")
    a = F.modl_iwae_step(params, x_u8, extra)
    assert a[-1] == 1
    monkeypatch.setenv("VAEMDL_FUSED", "0")
    b = F.modl_iwae_step(params, x_u8, extra)
    assert b[-1] == 3
    for u, v in zip(a[:3] + a[4:5], b[:3] + b[4:5]):
        assert torch.equal(u, v)
    if M == 5:      # both routes hand the per-pixel mixture sums from the forward to the backward pass: same arithmetic
        assert torch.equal(a[5], b[5])
    else:           # n_mix 10 as three launches keeps the two-pass gradient kernel: round-off apart
        assert relnorm(a[5], b[5]) <= 2e-6
    assert abs(a[3].item() - b[3].item()) <= 1e-6 * abs(b[3].item())
    assert torch.equal(default[5], b[5]) and torch.equal(default[0], b[0])
    ll64, g_ll, dp = a[0], a[4], a[5]
    lw = ll64 + extra.double()
    assert relnorm(g_ll, -torch.softmax(lw, 0) / B) < 1e-5
    for s_i, b_i in [(0, 0), (S - 1, B - 1), (2, B // 3)]:
        p64 = params[s_i, b_i].cpu().double()[None].requires_grad_(True)
        x64 = O.normalize_u8(x_u8[b_i].cpu(), torch.float64)[None]
        want = O.modl_log_prob(p64, x64).sum()
        assert abs(ll64[s_i, b_i].item() - want.item()) <= LL_RTOL * abs(want.item())
        (want * g_ll[s_i, b_i].item()).backward()
        assert relnorm(dp[s_i, b_i], p64.grad[0]) <= GRAD_RTOL


def test_config5_shard_bf16_parameters(F):
    """The config-5 shard with bfloat16 parameters: bit-identical sums to the float32 kernels on the widened tensor, the
    gradient is the float32 gradient rounded once."""
    S, B, H, W, M = 16, 32, 64, 64, 10
    gen = torch.Generator(device=DEV).manual_seed(55)
    pb = torch.randn(S, B, H, W, 10 * M, device=DEV, generator=gen).bfloat16()
    x_u8 = torch.randint(0, 256, (B, H, W, 3), dtype=torch.uint8, device=DEV, generator=gen)
    g = torch.randn(S, B, device=DEV, generator=gen)
    wide = pb.float()
    assert torch.equal(F.modl_log_likelihood(pb, x_u8, dtype=torch.float64), F.modl_log_likelihood(wide, x_u8, dtype=torch.float64))
    dp = F.modl_backward(pb, x_u8, g_image=g)
    assert dp.dtype == torch.bfloat16 and torch.equal(dp, F.modl_backward(wide, x_u8, g_image=g).bfloat16())


@pytest.mark.parametrize("S,B,H,W,M,reps", [(5, 64, 32, 32, 10, 200), (5, 128, 32, 32, 5, 100), (16, 32, 64, 64, 10, 30),
                                             (16, 16, 64, 64, 30, 15)])
def test_soak_bitwise_reproducible_under_back_to_back_steps(F, monkeypatch, S, B, H, W, M, reps):
    """The same step enqueued back to back `reps` times without host synchronisation (one-launch cooperative kernel for the
    small shapes, three launches with and without the handed-over mixture sums for the large ones): every repetition must
    reproduce the first bit for bit -- a race between warps, grid barriers or reused workspace would show up here."""
    gen = torch.Generator(device=DEV).manual_seed(900 + M)
    params = torch.randn(S, B, H, W, 10 * M, device=DEV, generator=gen)
    x_u8 = torch.randint(0, 256, (B, H, W, 3), dtype=torch.uint8, device=DEV, generator=gen)
    extra = torch.randn(S, B, device=DEV, generator=gen)
    for fused in (None, "1") if H == 32 else (None,):   # small shapes: also the opt-in cooperative one-launch kernel
        if fused:
            monkeypatch.setenv("VAEMDL_FUSED", fused)
        else:
            monkeypatch.delenv("VAEMDL_FUSED", raising=False)
        first = F.modl_iwae_step(params, x_u8, extra)
        assert first[-1] == (1 if fused else 3)
        ref_ll, ref_g, ref_dp, ref_elbo = first[0].clone(), first[4].clone(), first[5].clone(), first[3].clone()
        bad = torch.zeros((), dtype=torch.int64, device=DEV)
        for _ in range(reps):
            out = F.modl_iwae_step(params, x_u8, extra)
            bad += (out[0] != ref_ll).sum() + (out[4] != ref_g).sum() + (out[5] != ref_dp).sum() + (out[3] != ref_elbo).sum()
        assert int(bad.item()) == 0
        assert not bool(torch.isnan(ref_dp).any())


@pytest.mark.parametrize("name,H,W,M", [("cfg5_64_m30", 64, 64, 30), ("cfg5_128_m10", 128, 128, 10),
                                        ("cfg5_128_m30", 128, 128, 30), ("tm_128_m32", 128, 128, 32)])
def test_config5_other_sweep_points_full_size(F, V, name, H, W, M):
    """The other three points of BASELINE configs[4] at their full per-GPU size (16 importance samples x 32 images):
    whole-image oracle spot checks of the log-likelihood (1e-5) and of the gradient (1e-4) taken from both ends and the
    middle of the tensor.  cfg5_128_m30 is the one shape whose element offsets exceed 2^31 (8.39 M pixel-samples x 300
    floats = 2.5 G elements, 10 GB of parameters + 10 GB of gradients): its last image (15, 31) starts at element
    2,511,667,200 -- the 64-bit indexing path of the tile kernel, the bulk copies and the partial-sum bookkeeping.  n_mix 30 takes
    the one-pass gradient at this size; tm_128_m32 (n_mix 32, 2.68 G elements) is the same check for the two-pass gradient kernel on
    tensor memory."""
    S, B = 16, 32
    gen = torch.Generator(device=DEV).manual_seed(700 + M + H)
    params = torch.randn(S, B, H, W, 10 * M, device=DEV, generator=gen)
    x_u8 = torch.randint(0, 256, (B, H, W, 3), dtype=torch.uint8, device=DEV, generator=gen)
    if name in ("cfg5_128_m30", "tm_128_m32"):
        assert (S * B - 1) * H * W * 10 * M > 2 ** 31
    # `extra` balances the importance weights of every image (random parameters would give one sample all the weight and
    # leave the other samples' gradients in the float32 denormals): every spot-checked image carries a real gradient
    ll_first = F.modl_log_likelihood(params, x_u8, dtype=torch.float64)
    extra = (ll_first.mean(0, keepdim=True) - ll_first).float() + torch.randn(S, B, device=DEV, generator=gen)
    ll64, log_w, lme_b, elbo, g_ll, dp, launches = F.modl_iwae_step(params, x_u8, extra)
    assert launches == 3 and torch.equal(ll64, ll_first)
    assert g_ll.abs().min().item() > 1e-8
    lw = ll64 + extra.double()
    want_lme = torch.logsumexp(lw, 0) - math.log(S)
    assert ((lme_b.double() - want_lme).abs() / want_lme.abs()).max().item() < 1e-6
    assert abs(elbo.item() - want_lme.mean().item()) <= 1e-6 * abs(want_lme.mean().item())
    assert relnorm(g_ll, -torch.softmax(lw, 0) / B) < 1e-5
    # the per-pixel entry point sums to the fused per-image values everywhere (every pixel of the tensor is visited once)
    lp = F.modl_log_prob(params, x_u8)
    assert ((lp.double().sum((-1, -2)) - ll64).abs() / ll64.abs()).max().item() < 1e-12
    del lp
    # the logit gradients of every pixel sum to zero, and no element is NaN / left unwritten
    assert dp[..., :M].sum(-1).abs().max().item() < 2e-5 * g_ll.abs().max().item()
    assert bool(torch.isfinite(dp).all())
    for s_i, b_i in [(0, 0), (S - 1, B - 1), (7, B // 2)]:
        p64 = params[s_i, b_i].cpu().double()[None].requires_grad_(True)
        x64 = O.normalize_u8(x_u8[b_i].cpu(), torch.float64)[None]
        want = O.modl_log_prob(p64, x64).sum()
        assert abs(ll64[s_i, b_i].item() - want.item()) <= LL_RTOL * abs(want.item()), (name, s_i, b_i)
        (want * g_ll[s_i, b_i].item()).backward()
        assert relnorm(dp[s_i, b_i], p64.grad[0]) <= GRAD_RTOL, (name, s_i, b_i)
    del dp, params
    torch.cuda.empty_cache()
