"""Property tests of the CPU oracle (hypothesis): invariances the reference formulas must have, on random small inputs.
They pin the restatement from directions the fixed cases in test_oracle.py do not (the reference itself ships
no assertions for this path)."""
import math

import numpy as np
import torch
from hypothesis import given, settings
from hypothesis import strategies as st

import oracle as O

SET = dict(max_examples=25, deadline=None)


def _inputs(seed, B, H, W, M, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    p = torch.randn(B, H, W, 10 * M, generator=g, dtype=torch.float64) * scale
    x = torch.randint(0, 256, (B, H, W, 3), generator=g).double() / 255.0
    return p, x, g


@settings(**SET)
@given(seed=st.integers(0, 10**6), M=st.integers(1, 6), c=st.floats(-20, 20))
def test_adding_a_constant_to_the_mixture_logits_changes_nothing(seed, M, c):
    p, x, _ = _inputs(seed, 2, 3, 3, M)
    q = p.clone()
    q[..., :M] += c
    assert torch.allclose(O.modl_log_prob(p, x), O.modl_log_prob(q, x), atol=1e-9)
    assert torch.allclose(O.mdl_plain_log_prob(p, x), O.mdl_plain_log_prob(q, x), atol=1e-9)


@settings(**SET)
@given(seed=st.integers(0, 10**6), M=st.integers(2, 6))
def test_permuting_the_mixture_components_changes_nothing(seed, M):
    p, x, g = _inputs(seed, 2, 3, 3, M)
    perm = torch.randperm(M, generator=g)
    blocks = p.reshape(2, 3, 3, 10, M)[..., perm].reshape(2, 3, 3, 10 * M)
    assert torch.allclose(O.modl_log_prob(p, x), O.modl_log_prob(blocks, x), atol=1e-9)
    assert torch.allclose(O.modl_openai_iwae_log_prob(p[None], x)[0], O.modl_openai_iwae_log_prob(blocks[None], x)[0], atol=1e-9)


@settings(**SET)
@given(seed=st.integers(0, 10**6), M=st.integers(1, 4))
def test_a_pixel_distribution_sums_to_one_over_all_256_cubed_colours_marginally(seed, M):
    """With zero coefficients the three channels are independent inside a component: the red marginal over the 256
    bins must sum to one (edge bins absorb the tails) -- the defining property of the discretisation."""
    p, _, _ = _inputs(seed, 1, 1, 1, M, scale=0.7)
    for c in range(3):
        p[..., M + 3 * M * c + 2 * M:M + 3 * M * (c + 1)] = 0.0        # coefficients -> tanh(0) = 0
    k = torch.arange(256, dtype=torch.float64) / 255.0
    tot = torch.zeros((), dtype=torch.float64)
    # p(r = k) = sum over g, b of p(r,g,b) = mixture of the red marginals; evaluate it through a 1-component trick:
    loc, ls, logits = O.mdl_plain_get_mixture_params(p)
    w = torch.softmax(logits[0, 0, 0], -1)
    for m in range(M):
        lp = O.dlogistic_log_prob((k * 2 - 1), loc[0, 0, 0, 0, m], ls[0, 0, 0, 0, m], -1.0, 1.0, 256.0)
        tot = tot + w[m] * torch.exp(lp).sum()
    assert abs(tot.item() - 1.0) < 2e-3     # the low-probability branch replaces tiny masses by a density estimate


@settings(**SET)
@given(seed=st.integers(0, 10**6), S=st.integers(1, 6), B=st.integers(1, 5), c=st.floats(-1e4, 1e4))
def test_logmeanexp_shift_and_bounds(seed, S, B, c):
    g = torch.Generator().manual_seed(seed)
    lw = torch.randn(S, B, generator=g, dtype=torch.float64) * 5
    out = O.logmeanexp(lw, 0)
    assert torch.allclose(O.logmeanexp(lw + c, 0), out + c, atol=1e-8)
    assert bool((out <= lw.max(0).values + 1e-12).all()) and bool((out >= lw.mean(0) - 1e-12).all())   # Jensen


@settings(**SET)
@given(seed=st.integers(0, 10**6), M=st.integers(1, 5))
def test_sampler_outputs_are_valid_and_deterministic_in_the_noise(seed, M):
    g = torch.Generator().manual_seed(seed)
    l = torch.randn(2, 3, 3, 10 * M, generator=g)
    um = torch.rand(2, 3, 3, M, generator=g) * (1 - 2e-5) + 1e-5
    ul = torch.rand(2, 3, 3, 3, generator=g) * (1 - 2e-5) + 1e-5
    x, idx = O.sample_from_discretized_mix_logistic(l, M, um, ul)
    x2, idx2 = O.sample_from_discretized_mix_logistic(l, M, um, ul)
    assert torch.equal(x, x2) and torch.equal(idx, idx2)
    assert x.min() >= -1 and x.max() <= 1 and idx.min() >= 0 and idx.max() < M
    q = O.quantise(x * 0.5 + 0.5)
    assert q.dtype == torch.uint8 and np.all(np.abs(q.numpy().astype(np.float64) / 255 - (x.numpy() * 0.5 + 0.5)) <= 0.5 / 255 + 1e-12)


@settings(**SET)
@given(seed=st.integers(0, 10**6), S=st.integers(1, 4), B=st.integers(1, 4))
def test_iwae_loss_with_one_sample_is_the_negative_elbo(seed, S, B):
    g = torch.Generator().manual_seed(seed)
    lp = torch.randn(S, B, 2, 2, 3, generator=g, dtype=torch.float64)
    lpz = torch.randn(S, B, generator=g, dtype=torch.float64)
    lq = torch.randn(S, B, generator=g, dtype=torch.float64)
    loss, met = O.iwae_loss(lp, lpz, lq, (B, 2, 2, 3))
    eloss, _ = O.elbo_loss(lp, lpz, lq)
    assert loss.item() <= eloss.item() + 1e-9          # the IWAE bound is at least as tight as the ELBO
    if S == 1:
        assert abs(loss.item() - eloss.item()) < 1e-9
    assert abs(met["bpd"].item() + met["iwae_elbo"].item() / (math.log(2.0) * 12)) < 1e-12
