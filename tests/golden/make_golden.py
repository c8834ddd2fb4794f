"""Generates the golden fixtures in this directory from the float64 oracle (``oracle/ref.py``).

The reference itself (TensorFlow + TensorFlow-Probability) cannot be imported in this image, so these vectors are
NOT outputs of the reference -- they pin the oracle (any later change to it shows up as a fixture mismatch) and give
the GPU tests byte-identical inputs on every box.  The vectors that DO come from the reference's code are the
refsrc_*.npz files next to these (make_reference_golden.py: same inputs, outputs of the reference's own modules).

    python tests/golden/make_golden.py        # rewrites the .npz files
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle as O  # noqa: E402


def trained_like(g, S, B, H, W, M):
    """loc ~ U(-1,1), raw log-scale ~ N(-3,1): narrow components -> edge / low-probability branches fire often."""
    p = torch.randn(S, B, H, W, 10 * M, generator=g)
    rest = p[..., M:].reshape(S, B, H, W, 3, 3 * M).clone()
    rest[..., :M] = torch.rand(S, B, H, W, 3, M, generator=g) * 2 - 1
    rest[..., M:2 * M] = torch.randn(S, B, H, W, 3, M, generator=g) - 3.0
    return torch.cat([p[..., :M], rest.reshape(S, B, H, W, 9 * M)], -1).contiguous()


def modl_fixture(name, seed, S, B, H, W, M, kind):
    g = torch.Generator().manual_seed(seed)
    x_u8 = torch.randint(0, 256, (B, H, W, 3), dtype=torch.uint8, generator=g)
    # make sure both edge values occur
    x_u8.view(-1)[::17] = 0
    x_u8.view(-1)[5::23] = 255
    params = torch.randn(S, B, H, W, 10 * M, generator=g) if kind == "randn" else trained_like(g, S, B, H, W, M)
    g_image = torch.randn(S, B, generator=g)
    extra = torch.randn(S, B, generator=g)
    p64 = params.double().requires_grad_(True)
    x64 = O.normalize_u8(x_u8, torch.float64)
    lp = O.modl_log_prob(p64, x64)
    ll = lp.sum((-1, -2, -3))
    (ll * g_image.double()).sum().backward()
    grad_fixed = p64.grad.clone()
    # full IWAE chain with comparable importance weights (exercises the softmax over samples)
    p64b = params.double().requires_grad_(True)
    lpb = O.modl_log_prob(p64b, x64)
    llb = lpb.sum((-1, -2, -3)).detach()
    extra_bal = extra.double() + (llb.mean(0, keepdim=True) - llb)
    loss, met = O.iwae_loss(lpb, extra_bal, torch.zeros_like(extra_bal), x64.shape)
    loss.backward()
    np.savez_compressed(
        os.path.join(HERE, name), params=params.numpy(), x_u8=x_u8.numpy(), g_image=g_image.numpy(),
        lp=lp.detach().numpy()[..., 0], ll=ll.detach().numpy(), grad_fixed=grad_fixed.numpy(),
        extra=extra_bal.float().numpy(), loss=np.float64(loss.item()), grad_iwae=p64b.grad.numpy(),
        lme=O.logmeanexp(llb + extra_bal.float().double(), 0).numpy())


def dl_fixture(name, seed, S, B, H, W):
    g = torch.Generator().manual_seed(seed)
    both = torch.randn(S, B, H, W, 6, generator=g)
    both[..., :3] = torch.rand(S, B, H, W, 3, generator=g)   # loc in [0,1] like tests/test_hierarchical_setup.py:73
    x_u8 = torch.randint(0, 256, (B, H, W, 3), dtype=torch.uint8, generator=g)
    x_u8.view(-1)[::13] = 0
    x_u8.view(-1)[3::19] = 255
    g_image = torch.randn(S, B, generator=g)
    loc = both[..., :3].double().requires_grad_(True)
    ls = both[..., 3:].double().requires_grad_(True)
    x64 = O.normalize_u8(x_u8, torch.float64)
    lp = O.dlogistic_log_prob(x64, loc, ls, 0.0, 1.0, 256.0)
    ll = lp.sum((-1, -2, -3))
    (ll * g_image.double()).sum().backward()
    np.savez_compressed(os.path.join(HERE, name), both=both.numpy(), x_u8=x_u8.numpy(), g_image=g_image.numpy(),
                        lp=lp.detach().numpy(), ll=ll.detach().numpy(), dloc=loc.grad.numpy(), dls=ls.grad.numpy())


def sample_fixture(name, seed, N, H, W, M):
    g = torch.Generator().manual_seed(seed)
    l = torch.randn(N, H, W, 10 * M, generator=g)
    u_mix = torch.rand(N, H, W, M, generator=g) * (1 - 2e-5) + 1e-5
    u_log = torch.rand(N, H, W, 3, generator=g) * (1 - 2e-5) + 1e-5
    u_log_all = torch.rand(N, H, W, 3, M, generator=g) * (1 - 2e-5) + 1e-5
    x, idx = O.sample_from_discretized_mix_logistic(l, M, u_mix, u_log)
    x01, idx2 = O.modl_sample_mdl(l, u_mix, u_log_all)
    np.savez_compressed(os.path.join(HERE, name), l=l.numpy(), u_mix=u_mix.numpy(), u_log=u_log.numpy(),
                        u_log_all=u_log_all.numpy(), x_openai=x.numpy(), idx=idx.numpy().astype(np.uint8),
                        q_openai=O.quantise(x * 0.5 + 0.5).numpy(), x01_mdl=x01.numpy(), q_mdl=O.quantise(x01).numpy())


def plain_and_latent_fixture(name, seed, S, B, H, W, M, D):
    """utils/mdl_plain.py log-prob + gradient, its sampler / mean, and the latent-side Normal terms of models/loss.py:28-34."""
    import torch.distributions as td
    g = torch.Generator().manual_seed(seed)
    x_u8 = torch.randint(0, 256, (B, H, W, 3), dtype=torch.uint8, generator=g)
    x_u8.view(-1)[::11] = 0
    x_u8.view(-1)[4::29] = 255
    params = trained_like(g, S, B, H, W, M)
    g_image = torch.randn(S, B, generator=g)
    p64 = params.double().requires_grad_(True)
    lp = O.mdl_plain_log_prob(p64, O.normalize_u8(x_u8, torch.float64))
    ll = lp.sum((-1, -2))
    (ll * g_image.double()).sum().backward()
    u_mix = torch.rand(S, B, H, W, M, generator=g) * (1 - 2e-5) + 1e-5
    u_log = torch.rand(S, B, H, W, 3, M, generator=g) * (1 - 2e-5) + 1e-5
    xs, idx = O.mdl_plain_sample(params, u_mix, u_log)
    xm, _ = O.mdl_plain_sample(params, u_mix, None)
    # latent terms: z [S,B,D], q(z|x) = N(q_loc[B,D], q_scale[B,D]), p(z) = N(0,1), beta = 0.7
    q_loc = torch.randn(B, D, generator=g)
    q_scale = torch.rand(B, D, generator=g) + 0.5
    z = q_loc + q_scale * torch.randn(S, B, D, generator=g)
    z64, ql64, qs64 = z.double().requires_grad_(True), q_loc.double().requires_grad_(True), q_scale.double().requires_grad_(True)
    lpz = td.Normal(0.0, 1.0).log_prob(z64).sum(-1)
    lqzx = td.Normal(ql64, qs64).log_prob(z64).sum(-1)
    extra = 0.7 * (lpz - lqzx)
    (extra * g_image.double()).sum().backward()
    np.savez_compressed(
        os.path.join(HERE, name), params=params.numpy(), x_u8=x_u8.numpy(), g_image=g_image.numpy(),
        lp=lp.detach().numpy(), ll=ll.detach().numpy(), grad=p64.grad.numpy(), u_mix=u_mix.numpy(), u_log=u_log.numpy(),
        idx=idx.numpy().astype(np.uint8), q_sample=O.quantise(xs).numpy(), x_sample=xs.numpy(), x_mean=xm.numpy(),
        z=z.numpy(), q_loc=q_loc.numpy(), q_scale=q_scale.numpy(), lpz=lpz.detach().numpy(), lqzx=lqzx.detach().numpy(),
        extra=extra.detach().numpy(), dz=z64.grad.numpy(), dq_loc=ql64.grad.numpy(), dq_scale=qs64.grad.numpy())


if __name__ == "__main__":
    plain_and_latent_fixture("plain_m5_latent.npz", 401, 3, 4, 8, 8, 5, 12)
    modl_fixture("modl_m10_randn.npz", 101, 3, 4, 8, 8, 10, "randn")
    modl_fixture("modl_m5_trained.npz", 102, 2, 3, 8, 8, 5, "trained")
    modl_fixture("modl_m30_randn.npz", 103, 2, 2, 8, 4, 30, "randn")
    modl_fixture("modl_m7_ragged.npz", 104, 2, 3, 5, 7, 7, "randn")
    dl_fixture("dl_small.npz", 201, 3, 4, 8, 8)
    sample_fixture("sample_m10.npz", 301, 4, 8, 8, 10)
    print("golden fixtures written to", HERE)
