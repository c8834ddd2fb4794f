"""bfloat16 parameters: the kernels read a bf16 decoder output, widen it (in place in shared memory, or -- n_mix
10 / 20 / 30 -- pair by pair as they read a tile that stays bfloat16), compute in float32 and write a bf16 gradient.  Checked against (1) the float32 kernels on the widened parameters -- bit-identical
where both run the same kernel family -- and (2) the float64 oracle on the widened parameters."""
import pytest
import torch

import oracle as O
from util import GRAD_RTOL, LL_RTOL, relnorm, trained_like

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
BF16_GRAD_RTOL = 2.0 ** -8   # one round-to-nearest-even to 8 significant bits at the store (2^-9 per element, normwise < 2^-8)


@pytest.fixture(scope="module")
def F(built_lib):
    from vae_mdl_b200 import functional
    return functional


@pytest.fixture(scope="module")
def V(built_lib):
    import vae_mdl_b200
    return vae_mdl_b200


CASES = [(2, 3, 8, 8, 10, True), (3, 2, 8, 8, 5, True), (2, 2, 8, 8, 20, True), (2, 2, 8, 4, 30, True),      # tile kernels
         (1, 3, 3, 3, 10, True), (2, 1, 5, 7, 5, True), (1, 1, 1, 1, 10, True),                              # ragged tiles
         (2, 2, 8, 8, 7, False), (1, 2, 8, 8, 16, False), (2, 1, 6, 6, 3, False), (1, 2, 5, 5, 11, False),    # run-time kernel
         (1, 2, 4, 4, 64, False), (2, 2, 32, 32, 10, True), (1, 2, 16, 16, 12, False)]


@pytest.mark.parametrize("S,B,H,W,M,same_kernel", CASES)
def test_bf16_parameters_forward_backward(F, S, B, H, W, M, same_kernel):
    params, x_u8, g = trained_like(3300 + 7 * M + H, S, B, H, W, M)
    x_u8[0, 0, 0] = torch.tensor([0, 255, 0], dtype=torch.uint8)
    pb = params.bfloat16()
    wide = pb.float()                                    # what the kernels compute on
    g_image = torch.randn(S, B, generator=g)
    p64 = wide.double().requires_grad_(True)
    x64 = O.normalize_u8(x_u8, torch.float64)
    lp64 = O.modl_log_prob(p64, x64)[..., 0]
    ll64 = lp64.sum((-1, -2))
    (ll64 * g_image.double()).sum().backward()
    pbd, wd, xd, gd = pb.to(DEV), wide.to(DEV), x_u8.to(DEV), g_image.to(DEV)
    ll_b = F.modl_log_likelihood(pbd, xd, dtype=torch.float64)
    lp_b = F.modl_log_prob(pbd, xd)
    dp_b = F.modl_backward(pbd, xd, g_image=gd)
    assert dp_b.dtype == torch.bfloat16 and dp_b.shape == pbd.shape
    # (1) float32 kernels on the widened parameters
    ll_f = F.modl_log_likelihood(wd, xd, dtype=torch.float64)
    lp_f = F.modl_log_prob(wd, xd)
    dp_f = F.modl_backward(wd, xd, g_image=gd)
    if same_kernel:
        assert torch.equal(ll_b, ll_f) and torch.equal(lp_b, lp_f)
        assert torch.equal(dp_b, dp_f.bfloat16()), "the bf16 gradient must be the float32 gradient rounded once"
    else:
        assert ((ll_b - ll_f).abs() / ll_f.abs()).max().item() <= 1e-6
        assert (lp_b - lp_f).abs().max().item() <= 5e-5
        assert relnorm(dp_b.float(), dp_f) <= BF16_GRAD_RTOL
    # (2) the float64 oracle on the widened parameters (elements within rounding of the 1e-5 branch threshold excluded by
    # the image-level sum tolerance: trained-like parameters rounded to bf16 rarely sit there)
    err = ((ll_b.cpu() - ll64.detach()).abs() / ll64.detach().abs())
    assert err.median().item() <= LL_RTOL
    assert relnorm(dp_b.float().cpu(), p64.grad) <= 2 * BF16_GRAD_RTOL


def test_bf16_iwae_step_and_class_surface(F, V):
    S, B, H, W, M = 5, 6, 16, 16, 10
    g = torch.Generator().manual_seed(99)
    params = torch.randn(S, B, H, W, 10 * M, generator=g)
    x_u8 = torch.randint(0, 256, (B, H, W, 3), dtype=torch.uint8, generator=g)
    pb = params.bfloat16()
    p64 = pb.double().requires_grad_(True)
    x64 = O.normalize_u8(x_u8, torch.float64)
    lp = O.modl_log_prob(p64, x64)
    ll64 = lp.sum((-1, -2, -3)).detach()
    extra = (ll64.mean(0, keepdim=True) - ll64).float()
    loss64, _ = O.iwae_loss(lp, extra.double(), torch.zeros_like(extra).double(), x64.shape)
    loss64.backward()
    loss, lpxz, dp = V.modl_iwae_step(pb.to(DEV), x_u8.to(DEV), extra.to(DEV))
    assert dp.dtype == torch.bfloat16
    assert abs(loss.item() - loss64.item()) <= LL_RTOL * abs(loss64.item())
    assert ((lpxz.cpu() - ll64).abs() / ll64.abs()).max().item() <= LL_RTOL
    assert relnorm(dp.float().cpu(), p64.grad) <= 2 * BF16_GRAD_RTOL
    # the reference-shaped classes take the bf16 tensor as it is; autograd hands back a bf16 gradient
    leaf = pb.to(DEV).requires_grad_(True)
    x01 = V.normalize(x_u8.to(DEV))
    for cls in (V.MixtureDiscretizedLogistic, V.MixtureDiscretizedLogisticOpenaiIWAE):
        leaf.grad = None
        out = cls(leaf).log_prob(x01)
        assert out.dtype == torch.float32 and list(out.shape) == [S, B, H, W, 1]
        assert (out.detach().cpu().double() - lp.detach()).abs().max().item() < 5e-5
        out.sum().backward()
        assert leaf.grad is not None and leaf.grad.dtype == torch.bfloat16
    with pytest.raises(ValueError):
        V.PixelMixtureDiscretizedLogistic(leaf.detach()).log_prob(x01)


def test_bf16_no_write_outside_the_gradient_buffer(built_lib):
    """Guard band around the bf16 gradient (half the bytes of the float32 one) on ragged and full tiles."""
    L = built_lib
    import ctypes
    for (S, B, H, W, M) in [(2, 3, 8, 8, 10), (1, 3, 3, 3, 10), (2, 1, 5, 7, 5), (1, 2, 5, 5, 11), (1, 2, 8, 8, 16)]:
        n_img = S * B
        g = torch.Generator().manual_seed(S + B + H + M)
        pb = torch.randn(S, B, H, W, 10 * M, generator=g).bfloat16().to(DEV)
        x = torch.randint(0, 256, (B, H, W, 3), dtype=torch.uint8, generator=g).to(DEV)
        gi = torch.randn(S, B, generator=g).to(DEV)
        G = 64
        n = pb.numel()
        buf = torch.full((n + 2 * G,), -7.0, dtype=torch.bfloat16, device=DEV)
        dp = buf[G:G + n]
        if dp.data_ptr() % 16:
            continue
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        rc = L.vaemdl_modl_bwd_bf16(pb.data_ptr(), x.data_ptr(), 1, 0, 0, n_img, B, H, W, M, gi.data_ptr(), None, dp.data_ptr(), st)
        assert rc == 0
        torch.cuda.synchronize()
        assert bool((buf[:G] == -7.0).all()) and bool((buf[-G:] == -7.0).all())
        assert not bool(torch.isnan(dp.float()).any()) and bool((dp != -7.0).any())


@pytest.mark.parametrize("S,B,H,W,M", [(2, 3, 8, 8, 10), (2, 2, 8, 8, 20), (2, 2, 8, 4, 30), (1, 3, 3, 3, 10), (1, 1, 1, 1, 10),
                                       (3, 5, 32, 32, 10), (2, 3, 16, 16, 30), (4, 2, 64, 64, 10)])
def test_bf16_direct_route_with_forward_sums(F, monkeypatch, S, B, H, W, M):
    """n_mix 10 / 20 / 30 with the per-pixel sums handed from the forward to the backward call (vaemdl_modl_iwae_fwd_stats_bf16 /
    vaemdl_modl_bwd_stats_bf16): the tile stays bfloat16 in shared memory, two slots per warp.  Same arithmetic as the float32
    kernels on the widened parameters: the per-image sums are bit-identical, the gradient is the float32 one-pass gradient
    rounded once; both within tolerance of the float64 oracle; and identical to the widen-in-place route's sums."""
    params, x_u8, g = trained_like(5500 + 7 * M + H, S, B, H, W, M)
    pb = params.bfloat16()
    wide = pb.float()
    p64 = wide.double().requires_grad_(True)
    x64 = O.normalize_u8(x_u8, torch.float64)
    ll64 = O.modl_log_prob(p64, x64)[..., 0].sum((-1, -2))
    extra = (ll64.detach().mean(0, keepdim=True) - ll64.detach()).float() + torch.randn(S, B, generator=g)
    loss64 = -O.logmeanexp(ll64 + extra.double(), 0).mean()
    loss64.backward()
    pbd, wd, xd, ed = pb.to(DEV), wide.to(DEV), x_u8.to(DEV), extra.to(DEV)
    out_b = F.modl_iwae_forward(pbd, xd, ed, want_stats=True)
    assert out_b[5] is not None
    dp_b = F.modl_backward(pbd, xd, g_image=out_b[4], pix_stats=out_b[5])
    assert dp_b.dtype == torch.bfloat16
    # float32 kernels on the widened parameters, one-pass gradient from the same kind of sums
    monkeypatch.setenv("VAEMDL_STATS", "all")
    out_f = F.modl_iwae_forward(wd, xd, ed, want_stats=True)
    dp_f = F.modl_backward(wd, xd, g_image=out_f[4], pix_stats=out_f[5])
    monkeypatch.delenv("VAEMDL_STATS")
    assert torch.equal(out_b[0], out_f[0]) and torch.equal(out_b[4], out_f[4]) and torch.equal(out_b[5], out_f[5])
    assert torch.equal(dp_b, dp_f.bfloat16()), "the bf16 gradient must be the float32 one-pass gradient rounded once"
    # the widen-in-place route (no sums handed over) gives the same forward results
    monkeypatch.setenv("VAEMDL_BF16_WIDEN", "1")
    out_w = F.modl_iwae_forward(pbd, xd, ed)
    monkeypatch.delenv("VAEMDL_BF16_WIDEN")
    assert torch.equal(out_b[0], out_w[0])
    # float64 oracle
    assert abs(out_b[3].item() + loss64.item()) <= LL_RTOL * abs(loss64.item())
    assert relnorm(dp_b.float().cpu(), p64.grad) <= 2 * BF16_GRAD_RTOL
    # the one-call step takes the same route
    step = F.modl_iwae_step(pbd, xd, ed)
    assert torch.equal(step[5], dp_b) and torch.equal(step[0], out_b[0])
