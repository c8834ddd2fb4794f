"""GPU tests of the host-buffer step (vaemdl_modl_iwae_step_host): same numbers as the device-resident path."""
import pytest
import torch

import oracle as O
from util import LL_RTOL, assert_grad_close, canonical

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("S,B,H,W,M,chunk", [(3, 10, 8, 8, 10, 3), (5, 7, 32, 32, 5, 0), (2, 4, 8, 8, 7, 1), (2, 5, 16, 16, 30, 2)])
def test_host_step_matches_device_step_and_oracle(built_lib, S, B, H, W, M, chunk):
    import vae_mdl_b200 as V
    params, x_u8, g = canonical(300 + M, S, B, H, W, M)
    extra = torch.randn(S, B, generator=g)
    # comparable importance weights so the softmax over samples matters
    ll64 = O.modl_log_prob(params.double(), O.normalize_u8(x_u8, torch.float64)).sum((-1, -2, -3))
    extra = (extra.double() + ll64.mean(0, keepdim=True) - ll64).float()
    ph, xh, eh = params.pin_memory(), x_u8.pin_memory(), extra.pin_memory()
    dp = torch.empty_like(params).pin_memory()
    ll = torch.empty(S, B).pin_memory()
    lme = torch.empty(B).pin_memory()
    elbo = torch.empty(1).pin_memory()
    rc = built_lib.vaemdl_modl_iwae_step_host(ph.data_ptr(), xh.data_ptr(), eh.data_ptr(), S, B, H, W, M, dp.data_ptr(),
                                              ll.data_ptr(), lme.data_ptr(), elbo.data_ptr(), chunk)
    assert rc == 0
    loss_d, lpxz_d, dp_d = V.modl_iwae_step(params.to(DEV), x_u8.to(DEV), extra.to(DEV))
    # not bitwise: which of the two (equally accurate) exp(-h) evaluations a pixel gets is a per-warp-tile decision,
    # and chunking the batch changes which pixels share a tile
    assert torch.allclose(ll, lpxz_d.float().cpu(), rtol=1e-6, atol=0)
    assert torch.allclose(dp, dp_d.cpu(), rtol=1e-4, atol=1e-7)
    assert abs(-elbo.item() - loss_d.item()) <= 1e-6 * abs(loss_d.item())
    # and against the oracle
    p64 = params.double().requires_grad_(True)
    x64 = O.normalize_u8(x_u8, torch.float64)
    loss64, _ = O.iwae_loss(O.modl_log_prob(p64, x64), extra.double(), torch.zeros(S, B, dtype=torch.float64), x64.shape)
    loss64.backward()
    assert abs(-elbo.item() - loss64.item()) <= LL_RTOL * abs(loss64.item())
    assert_grad_close(dp, p64.grad, M)
    # forward only (evaluation): no gradient buffer
    rc = built_lib.vaemdl_modl_iwae_step_host(ph.data_ptr(), xh.data_ptr(), None, S, B, H, W, M, None, ll.data_ptr(),
                                              lme.data_ptr(), elbo.data_ptr(), chunk)
    assert rc == 0
    want = O.logmeanexp(ll64, 0)
    assert ((lme.double() - want).abs() / want.abs()).max().item() <= LL_RTOL
    built_lib.vaemdl_host_release()
