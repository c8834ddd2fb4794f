"""The two-pass gradient kernel that keeps its tile in tensor memory (csrc/modl_tm.cuh) against (1) the shared-memory kernel it
replaces -- same arithmetic in the same order, so the gradients must be bit-identical -- and (2) the float64 oracle."""
import pytest
import torch

import oracle as O
from util import GRAD_RTOL, relnorm, trained_like

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def F(built_lib):
    from vae_mdl_b200 import functional
    return functional


# (S, B, H, W, M): every n_mix with a tensor-memory instantiation (MC in 6 ... 14, 1 ... 8 lanes per pixel), full and ragged tiles
CASES = [(2, 3, 8, 8, 10), (5, 4, 32, 32, 10), (1, 1, 8, 4, 10), (2, 2, 8, 8, 20), (3, 2, 16, 16, 20), (2, 3, 10, 8, 30),
         (2, 2, 8, 8, 6), (2, 2, 8, 8, 8), (1, 2, 8, 8, 12), (1, 2, 8, 8, 14), (1, 2, 8, 8, 16), (1, 2, 10, 8, 18), (1, 1, 8, 8, 24),
         (1, 1, 8, 8, 28), (1, 2, 8, 8, 32), (1, 1, 10, 8, 36), (1, 1, 8, 8, 40), (1, 1, 8, 8, 48), (1, 1, 12, 8, 50), (1, 1, 8, 8, 56),
         (1, 1, 12, 8, 60), (1, 1, 8, 8, 64),
         (1, 3, 3, 3, 10), (2, 1, 5, 7, 20), (1, 1, 1, 1, 10), (1, 2, 7, 5, 30), (3, 5, 9, 7, 10), (1, 2, 5, 5, 16), (2, 2, 8, 8, 18)]


@pytest.mark.parametrize("S,B,H,W,M", CASES)
def test_tm_gradient_is_bit_identical_and_matches_oracle(F, monkeypatch, S, B, H, W, M):
    params, x_u8, g = trained_like(5100 + 3 * M + W, S, B, H, W, M)
    x_u8[0, 0, 0] = torch.tensor([0, 255, 17], dtype=torch.uint8)   # both edge bins
    if H > 1:
        params[0, 0, 1, 1, :M] = -200.0                              # a pixel whose mixture sum leaves the float32 range:
        params[0, 0, 1, 1, M:2 * M] = 40.0                           # log-domain fallback inside the second pass
    g_image = torch.randn(S, B, generator=g)
    g_pixel = torch.randn(S, B, H, W, generator=g)
    pd, xd, gid, gpd = params.to(DEV), x_u8.to(DEV), g_image.to(DEV), g_pixel.to(DEV)
    monkeypatch.setenv("VAEMDL_STATS", "none")   # the two-pass kernels for every n_mix
    monkeypatch.setenv("VAEMDL_TM", "0")
    ref_i = F.modl_backward(pd, xd, g_image=gid)
    ref_ip = F.modl_backward(pd, xd, g_image=gid, g_pixel=gpd)
    monkeypatch.setenv("VAEMDL_TM", "1")
    tm_i = F.modl_backward(pd, xd, g_image=gid)
    tm_ip = F.modl_backward(pd, xd, g_image=gid, g_pixel=gpd)
    torch.cuda.synchronize()
    assert torch.equal(tm_i, ref_i)
    assert torch.equal(tm_ip, ref_ip)
    p64 = params.double().requires_grad_(True)
    x64 = O.normalize_u8(x_u8, torch.float64)
    ll64 = O.modl_log_prob(p64, x64)[..., 0].sum((-1, -2))
    (ll64 * g_image.double()).sum().backward()
    assert relnorm(tm_i.cpu(), p64.grad) <= GRAD_RTOL


def test_tm_plain_mixture_bit_identical(F, monkeypatch):
    S, B, H, W, M = 2, 2, 8, 8, 10
    params, x_u8, g = trained_like(5207, S, B, H, W, M)
    gid = torch.randn(S, B, generator=g).to(DEV)
    pd, xd = params.to(DEV), x_u8.to(DEV)
    monkeypatch.setenv("VAEMDL_TM", "0")
    ref = F.modl_backward(pd, xd, g_image=gid, plain=True)
    monkeypatch.setenv("VAEMDL_TM", "1")
    tm = F.modl_backward(pd, xd, g_image=gid, plain=True)
    assert torch.equal(tm, ref)


def test_tm_many_tiles_per_warp_and_guard_band(F, monkeypatch):
    """A run of several tiles per warp (the swap of consecutive tiles through tensor memory) at a size where every warp of the
    grid has work; nothing is written outside the gradient buffer."""
    S, B, H, W, M = 8, 16, 64, 64, 10    # 524,288 pixel-samples: 16,384 tiles, ~7 per warp
    g = torch.Generator().manual_seed(77)
    params = torch.randn(S, B, H, W, 10 * M, generator=g)
    params[..., 2 * M:3 * M] -= 3.0
    x_u8 = torch.randint(0, 256, (B, H, W, 3), dtype=torch.uint8, generator=g)
    gid = torch.randn(S, B, generator=g).to(DEV)
    pd, xd = params.to(DEV), x_u8.to(DEV)
    monkeypatch.setenv("VAEMDL_TM", "0")
    ref = F.modl_backward(pd, xd, g_image=gid)
    monkeypatch.setenv("VAEMDL_TM", "1")
    tm = F.modl_backward(pd, xd, g_image=gid)
    assert torch.equal(tm, ref)
    guard = torch.full((pd.numel() + 8192,), 7.25, device=DEV)
    dp = guard[4096:4096 + pd.numel()].view_as(pd)
    from vae_mdl_b200 import _abi
    L = _abi.lib()
    st = _abi.stream_ptr(pd.device)
    rc = L.vaemdl_modl_bwd(pd.data_ptr(), xd.data_ptr(), 1, 0, 0, S * B, B, H, W, M, gid.data_ptr(), None, dp.data_ptr(), st)
    torch.cuda.synchronize()
    assert rc == 0 and torch.equal(dp, ref)
    assert bool((guard[:4096] == 7.25).all()) and bool((guard[4096 + pd.numel():] == 7.25).all())

