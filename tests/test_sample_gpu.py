"""GPU tests of the explicit-noise samplers: mixture indices and quantised values bit-exact against the float64 oracle."""
import numpy as np
import pytest
import torch

import oracle as O
from util import golden

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def V(built_lib):
    import vae_mdl_b200
    return vae_mdl_b200


def test_golden_openai_variant_bit_exact(V):
    z = golden("sample_m10")
    l = torch.from_numpy(z["l"]).to(DEV)
    x, xq, idx = V.sample_from_discretized_mix_logistic(l, 10, torch.from_numpy(z["u_mix"]).to(DEV),
                                                        torch.from_numpy(z["u_log"]).to(DEV), return_index=True,
                                                        return_quantised=True)
    assert np.array_equal(idx.cpu().numpy(), z["idx"])
    assert np.array_equal(xq.cpu().numpy(), z["q_openai"])
    assert np.abs(x.cpu().double().numpy() - z["x_openai"]).max() < 1e-6
    assert x.min().item() >= -1.0 and x.max().item() <= 1.0


def test_golden_mdl_variant_bit_exact(V):
    z = golden("sample_m10")
    d = V.MixtureDiscretizedLogistic(torch.from_numpy(z["l"]).to(DEV))
    x, xq, idx = d.sample(1, u_mix=torch.from_numpy(z["u_mix"]).to(DEV)[None], u_log=torch.from_numpy(z["u_log_all"]).to(DEV)[None],
                          return_index=True, return_quantised=True)
    assert x.shape == (1, 4, 8, 8, 3)
    assert np.array_equal(idx[0].cpu().numpy(), z["idx"])
    assert np.array_equal(xq[0].cpu().numpy(), z["q_mdl"])
    assert np.abs(x[0].cpu().double().numpy() - z["x01_mdl"]).max() < 1e-6


@pytest.mark.parametrize("M", [5, 10, 30, 7])
def test_seeded_bit_exact_counts(V, M):
    """BASELINE config 3 at reduced size: count index / quantised mismatches (must be zero)."""
    g = torch.Generator().manual_seed(40 + M)
    N, H, W = 24, 32, 32
    l = torch.randn(N, H, W, 10 * M, generator=g)
    u_mix = torch.rand(N, H, W, M, generator=g) * (1 - 2e-5) + 1e-5
    u_log = torch.rand(N, H, W, 3, generator=g) * (1 - 2e-5) + 1e-5
    x64, idx64 = O.sample_from_discretized_mix_logistic(l, M, u_mix, u_log)
    x, xq, idx = V.sample_from_discretized_mix_logistic(l.to(DEV), M, u_mix.to(DEV), u_log.to(DEV), return_index=True,
                                                        return_quantised=True)
    assert int((idx.cpu().long() != idx64).sum()) == 0
    assert int((xq.cpu() != O.quantise(x64 * 0.5 + 0.5)).sum()) == 0
    assert (x.cpu().double() - x64).abs().max().item() < 1e-6


def test_class_sample_shapes_and_tiling(V):
    g = torch.Generator().manual_seed(9)
    S, B, H, W, M = 2, 3, 8, 8, 5
    l = torch.randn(S, B, H, W, 10 * M, generator=g)
    n = 4
    u_mix = torch.rand(n, S, B, H, W, M, generator=g) * (1 - 2e-5) + 1e-5
    u_log = torch.rand(n, S, B, H, W, 3, generator=g) * (1 - 2e-5) + 1e-5
    # IWAE wrapper: [n, S, B, H, W, 3] in [0,1]; the parameters are re-used for every n without being tiled
    d = V.MixtureDiscretizedLogisticOpenaiIWAE(l.to(DEV))
    x = d.sample(n, u_mix=u_mix.to(DEV), u_log=u_log.to(DEV))
    want, _ = O.modl_openai_iwae_sample(l, n, u_mix.reshape(n * S * B, H, W, M), u_log.reshape(n * S * B, H, W, 3))
    assert x.shape == (n, S, B, H, W, 3)
    assert (x.cpu().double() - want).abs().max().item() < 1e-6
    assert d.sample().shape == (S, B, H, W, 3) and d.sample([2]).shape == (2, S, B, H, W, 3)
    assert d.mean(n=8).shape == (S, B, H, W, 3)
    # 4-D wrapper: [-1,1]
    d4 = V.MixtureDiscretizedLogisticOpenai(l[0].to(DEV))
    x4 = d4.sample(n, u_mix=u_mix[:, 0].to(DEV), u_log=u_log[:, 0].to(DEV))
    want4, _ = O.modl_openai_sample(l[0], n, u_mix[:, 0].reshape(n * B, H, W, M), u_log[:, 0].reshape(n * B, H, W, 3))
    assert x4.shape == (n, B, H, W, 3)
    assert (x4.cpu().double() - want4).abs().max().item() < 1e-6
    # utils/mdl.py class: sample() drops the leading dim (tfd semantics), values in [0,1]
    dm = V.MixtureDiscretizedLogistic(l.to(DEV))
    s0 = dm.sample()
    assert s0.shape == (S, B, H, W, 3) and s0.min().item() >= 0 and s0.max().item() <= 1
    assert dm.mean(n=4).shape == (S, B, H, W, 3)
    # attributes of the reference class (utils/mdl.py:43-54)
    assert dm.n_mix == M and dm.shape == [S, B, H, W, 10 * M] and dm.axes == [-1, -2, -3]
    assert abs(dm.interval_width - 2 / 255) < 1e-15 and (dm.low, dm.high) == (-1.0, 1.0)
    dm.axes = [-1, -2]
    assert dm.axes == [-1, -2]


def test_random_sampler_statistics(V):
    """Device-drawn noise: the empirical mixture frequencies follow softmax(logits)."""
    M = 5
    logits = torch.tensor([0.0, 1.0, -1.0, 2.0, 0.5])
    l = torch.zeros(64, 32, 32, 10 * M)
    l[..., :M] = logits
    gen = torch.Generator(device=DEV).manual_seed(0)
    x, idx = V.sample_from_discretized_mix_logistic(l.to(DEV), M, generator=gen, return_index=True)
    freq = torch.bincount(idx.flatten().long().cpu(), minlength=M).double() / idx.numel()
    assert (freq - torch.softmax(logits.double(), 0)).abs().max().item() < 1e-2


def test_gumbel_near_ties_are_decided_in_float64(V):
    """Winners that lead by less than float32 can resolve (1e-9 .. 1e-5), exact ties (first maximum wins) and a
    degenerate u -> the kernel's float32 search must hand over to float64 and agree with the oracle everywhere."""
    g = torch.Generator().manual_seed(77)
    N, H, W, M = 6, 16, 16, 10
    l = torch.randn(N, H, W, 10 * M, generator=g)
    u_mix = torch.rand(N, H, W, M, generator=g) * (1 - 2e-5) + 1e-5
    u_log = torch.rand(N, H, W, 3, generator=g) * (1 - 2e-5) + 1e-5
    # make component 7 trail the current winner by a tiny margin that differs per image (float64 arithmetic)
    gum = l[..., :M].double() - torch.log(-torch.log(u_mix.double()))
    best, arg = gum.max(-1)
    margins = torch.tensor([0.0, 1e-9, 1e-7, 1e-6, 1e-5, -1e-7]).view(N, 1, 1)
    other = torch.where(arg == 7, torch.full_like(arg, 2), torch.full_like(arg, 7))
    noise_other = -torch.log(-torch.log(u_mix.double().gather(-1, other[..., None])))[..., 0]
    new_logit = (best - margins - noise_other).float()      # float32 logit: the realised margin is whatever float32 leaves
    l[..., :M].scatter_(-1, other[..., None], new_logit[..., None])
    x64, idx64 = O.sample_from_discretized_mix_logistic(l, M, u_mix, u_log)
    x, xq, idx = V.sample_from_discretized_mix_logistic(l.to(DEV), M, u_mix.to(DEV), u_log.to(DEV), return_index=True,
                                                        return_quantised=True)
    assert int((idx.cpu().long() != idx64).sum()) == 0
    assert int((xq.cpu() != O.quantise(x64 * 0.5 + 0.5)).sum()) == 0
    # exact ties: identical logits and identical noise in two components -> the first one wins on both sides
    l2 = torch.randn(2, 8, 8, 10 * M, generator=g)
    l2[..., 3] = l2[..., 8] = 9.0
    u2 = torch.rand(2, 8, 8, M, generator=g) * (1 - 2e-5) + 1e-5
    u2[..., 8] = u2[..., 3]
    ul2 = torch.rand(2, 8, 8, 3, generator=g) * (1 - 2e-5) + 1e-5
    _, i64 = O.sample_from_discretized_mix_logistic(l2, M, u2, ul2)
    _, i2 = V.sample_from_discretized_mix_logistic(l2.to(DEV), M, u2.to(DEV), ul2.to(DEV), return_index=True)
    assert int((i2.cpu().long() != i64).sum()) == 0 and int((i64 == 8).sum()) == 0


def test_sample_grid_is_the_quantised_sample_canvas(V):
    """models/model05.py:200-216 on bytes: an 8x8 grid of quantised samples straight from the sampling kernel."""
    g = torch.Generator().manual_seed(12)
    n, H, W, M = 4, 8, 8, 5
    l = torch.randn(n * n, H, W, 10 * M, generator=g)
    um = torch.rand(1, n * n, H, W, M, generator=g) * (1 - 2e-5) + 1e-5
    ul = torch.rand(1, n * n, H, W, 3, M, generator=g) * (1 - 2e-5) + 1e-5
    d = V.MixtureDiscretizedLogistic(l.to(DEV))
    canvas = V.sample_grid(d, n, sample_shape=1, u_mix=um.to(DEV), u_log=ul.to(DEV))
    x01, _ = O.modl_sample_mdl(l, um[0], ul[0])
    want = V.fill_canvas(O.quantise(x01), n, H, W, 3)
    assert canvas.dtype == torch.uint8 and canvas.shape == (n * H, n * W, 3)
    assert torch.equal(canvas.cpu(), want)
