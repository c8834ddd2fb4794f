"""Parity against outputs of the REFERENCE'S OWN SOURCE (tests/golden/refsrc_*.npz, made by
tests/golden/make_reference_golden.py: nbip/vae-mdl's unmodified modules executed over oracle/tf_shim).

CPU part: the oracle restatement (oracle/ref.py) agrees with what the reference's code computes, float64 and float32.
GPU part: the CUDA kernels, through the package's reference-shaped classes, agree with it at north_star's tolerances
(per-image log-likelihood / loss / BPD 1e-5 relative, gradients 1e-4, sampled indices / quantised bytes bit-exact).
"""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import oracle as O
from util import GOLDEN, GRAD_RTOL, LL_RTOL, assert_grad_close, assert_ll_close, golden, relnorm

DEV = "cuda:0"
MODL = [("modl_m10_randn", 10), ("modl_m5_trained", 5), ("modl_m30_randn", 30), ("modl_m7_ragged", 7)]

# The mdl.py / discretized_logistic.py classes add log(interval_width) ROUNDED TO float32 in their low-probability
# branch (utils/mdl.py:163, utils/discretized_logistic.py:33: `tf.cast(tf.math.log(...), tf.float32)`), the OpenAI
# functions add the float64 constant np.log(127.5) (utils/mdl_openai.py:148).  float32(log(2/255)) is 1.2e-7 away from
# the real number, so float64 runs of the two formulations differ by that much per low-probability sub-pixel.
CAST_ABS = 4e-7
PLAIN_BINS = (-0.95, 0.9, 64.0)   # tests/golden/make_reference_golden.py: the non-default PixelMixtureDiscretizedLogistic


def refsrc(name):
    return np.load(os.path.join(GOLDEN, "refsrc_" + name + ".npz"))


def t64(a):
    return torch.from_numpy(np.asarray(a)).double()


# ------------------------------------------------------------------------------------------------------------------
# CPU: the fixtures are what the reference computes; the oracle restates it
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="the reference is only mounted in the build container")
def test_fixtures_reproduce_from_the_reference_source():
    r = subprocess.run([sys.executable, os.path.join(GOLDEN, "make_reference_golden.py"), "--check"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "reproduce bit for bit" in r.stdout


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="the reference is only mounted in the build container")
def test_oracle_matches_reference_source_on_random_cases():
    """tests/golden/fuzz_reference.py: 40 random shapes / n_mix / scales / (low, high, levels), values and gradients, the
    oracle against the reference's own modules executed over oracle/tf_shim (in a subprocess: it puts the reference on
    sys.path)."""
    r = subprocess.run([sys.executable, os.path.join(GOLDEN, "fuzz_reference.py"), "40", "20261018"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr


@pytest.mark.parametrize("name,M", MODL)
def test_oracle_matches_reference_source_modl_f64(name, M):
    fx, rs = golden(name), refsrc(name)
    params = torch.from_numpy(fx["params"]).double().requires_grad_(True)
    x = O.normalize_u8(torch.from_numpy(fx["x_u8"]), torch.float32).double()
    g_image = t64(fx["g_image"])
    lp = O.modl_log_prob(params, x)
    ll = lp.sum((-1, -2, -3))
    (ll * g_image).sum().backward()
    n_sub = 3 * M
    assert (lp.detach()[..., 0] - t64(rs["mdl_lp_f64"])).abs().max().item() <= CAST_ABS * n_sub
    assert_ll_close(ll, rs["mdl_ll_f64"], rtol=1e-8)
    assert relnorm(params.grad, t64(rs["mdl_grad_fixed_f64"])) <= 1e-7        # the float32-rounded constant shifts responsibilities
    # the IWAE / OpenAI formulation: exact constants on both sides -> round-off only
    p2 = torch.from_numpy(fx["params"]).double().requires_grad_(True)
    lp2 = O.modl_openai_iwae_log_prob(p2, x)
    (lp2.sum((-1, -2, -3)) * g_image).sum().backward()
    assert (lp2.detach()[..., 0] - t64(rs["iwae_cls_lp_f64"])).abs().max().item() <= 1e-11
    assert relnorm(p2.grad, t64(rs["iwae_cls_grad_fixed_f64"])) <= 1e-12
    lp3 = O.modl_openai_log_prob(p2.detach()[0], x * 2.0 - 1.0)
    assert (lp3 - t64(rs["openai_cls_lp_s0_f64"])).abs().max().item() <= 1e-11
    tot = O.discretized_mix_logistic_loss(x * 2.0 - 1.0, p2.detach()[0], sum_all=True)
    assert abs(tot.item() - float(rs["openai_sum_all_s0_f64"])) <= 1e-12 * abs(tot.item())
    # models/loss.py::iwae_loss end to end
    p4 = torch.from_numpy(fx["params"]).double().requires_grad_(True)
    extra = t64(fx["extra"])
    loss, met = O.iwae_loss(O.modl_log_prob(p4, x), extra, torch.zeros_like(extra), x.shape)
    loss.backward()
    assert abs(loss.item() - float(rs["iwae_loss_f64"])) <= 1e-8 * abs(loss.item())
    assert relnorm(p4.grad, t64(rs["iwae_grad_f64"])) <= 1e-6
    assert abs(met["bpd"].item() - float(rs["iwae_bpd_f64"])) <= 1e-6 * abs(met["bpd"].item())   # reference casts n_dims / log 2 to float32
    assert relnorm(met["kl"], t64(rs["iwae_kl_f64"])) <= 1e-12
    eloss, _ = O.elbo_loss(O.modl_log_prob(p4.detach(), x), extra, torch.zeros_like(extra))
    assert abs(eloss.item() - float(rs["elbo_loss_f64"])) <= 1e-8 * abs(eloss.item())
    lw = t64(rs["iwae_lpxz_f64"]) + extra
    assert (O.logmeanexp(lw, 0) - t64(rs["logmeanexp_axis0_f64"])).abs().max().item() <= 1e-11


@pytest.mark.parametrize("name,M", MODL)
def test_oracle_ref32_tracks_reference_source_f32(name, M):
    """The float32 flavour of the oracle (the CPU baseline bench.py times) against the reference's code run in float32,
    and how far the reference's OWN float32 result is from its float64 one -- the head-room the tolerances leave."""
    fx, rs = golden(name), refsrc(name)
    params = torch.from_numpy(fx["params"]).requires_grad_(True)
    x = O.normalize_u8(torch.from_numpy(fx["x_u8"]), torch.float32)
    lp = O.modl_log_prob(params, x)
    ll = lp.sum((-1, -2, -3))
    (ll * torch.from_numpy(fx["g_image"])).sum().backward()
    assert (lp.detach()[..., 0] - torch.from_numpy(rs["mdl_lp_f32"])).abs().max().item() <= 1e-5
    assert_ll_close(ll, rs["mdl_ll_f32"], rtol=1e-6)
    assert relnorm(params.grad, torch.from_numpy(rs["mdl_grad_fixed_f32"])) <= 1e-5
    # float32 reference vs float64 reference: per-image sums are inside 1e-5, the IWAE gradient is NOT always inside 1e-4
    assert_ll_close(torch.from_numpy(rs["mdl_ll_f32"]), rs["mdl_ll_f64"], rtol=LL_RTOL)


def test_oracle_matches_reference_source_dl_and_model06():
    fx, rs = golden("dl_small"), refsrc("dl_small")
    both = torch.from_numpy(fx["both"]).double()
    x = O.normalize_u8(torch.from_numpy(fx["x_u8"]), torch.float32).double()
    loc, ls = both[..., :3].clone().requires_grad_(True), both[..., 3:].clone().requires_grad_(True)
    lp = O.dlogistic_log_prob(x, loc, ls, 0.0, 1.0, 256.0)
    ll = lp.sum((-1, -2, -3))
    (ll * t64(fx["g_image"])).sum().backward()
    assert (lp.detach() - t64(rs["lp_f64"])).abs().max().item() <= CAST_ABS
    assert_ll_close(ll, rs["ll_f64"], rtol=1e-8)
    assert relnorm(loc.grad, t64(rs["dloc_f64"])) <= 1e-10 and relnorm(ls.grad, t64(rs["dls_f64"])) <= 1e-10
    # models/model06.py::loss_fn
    import torch.distributions as td
    z1, z2 = t64(rs["m6_z1"]), t64(rs["m6_z2"])
    n = lambda k: td.Normal(t64(rs["m6_%s_loc" % k]), t64(rs["m6_%s_scale" % k]))
    lqz2z1 = n("q2").log_prob(z2).sum(-1)
    lqz1x = n("q1").log_prob(z1).sum(-1)
    lpz2 = td.Normal(0.0, 1.0).log_prob(z2).sum(-1)
    lpz1z2 = n("p1").log_prob(z1).sum(-1)
    loc2, ls2 = both[..., :3].clone().requires_grad_(True), both[..., 3:].clone().requires_grad_(True)
    loss, met = O.model06_loss(O.dlogistic_log_prob(x, loc2, ls2, 0.0, 1.0, 256.0), lpz2, lqz2z1, lpz1z2, lqz1x, x.shape)
    loss.backward()
    assert abs(loss.item() - float(rs["m6_loss_f64"])) <= 1e-8 * abs(loss.item())
    assert relnorm(loc2.grad, t64(rs["m6_dloc_f64"])) <= 1e-8 and relnorm(ls2.grad, t64(rs["m6_dls_f64"])) <= 1e-8
    for k, v in (("lqz1x", lqz1x), ("lqz2z1", lqz2z1), ("lpz2", lpz2), ("lpz1z2", lpz1z2), ("kl1", met["kl1"]), ("kl2", met["kl2"])):
        assert relnorm(v, t64(rs["m6_%s_f64" % k])) <= 1e-12, k
    assert abs(met["bpd"].item() - float(rs["m6_bpd_f64"])) <= 1e-6 * abs(met["bpd"].item())


def test_oracle_matches_reference_source_samplers():
    fx, rs = golden("sample_m10"), refsrc("sample_m10")
    l = torch.from_numpy(fx["l"])
    u_mix, u_log, u_all = (torch.from_numpy(fx[k]) for k in ("u_mix", "u_log", "u_log_all"))
    x, idx = O.sample_from_discretized_mix_logistic(l, 10, u_mix, u_log)
    assert (x - t64(rs["x_openai_f64"])).abs().max().item() <= 1e-12
    assert (x - t64(rs["x_openai_cls_f64"])[0]).abs().max().item() <= 1e-12
    assert (x * 0.5 + 0.5 - t64(rs["x01_iwae_cls_f64"])).abs().max().item() <= 1e-12
    x01, _ = O.modl_sample_mdl(l, u_mix, u_all)
    assert (x01 - t64(rs["x01_mdl_f64"])).abs().max().item() <= 1e-12
    xd = O.dlogistic_sample(l[..., :3], l[..., 3:6], u_log, 0.0, 1.0)
    assert (xd - t64(rs["x_dl_f64"])).abs().max().item() <= 1e-12
    # the quantised bytes of the float64 and of the float32 run of the reference's code: identical here
    assert torch.equal(O.quantise(t64(rs["x_openai_f64"]) * 0.5 + 0.5), torch.from_numpy(fx["q_openai"]))
    assert torch.equal(O.quantise(t64(rs["x01_mdl_f64"])), torch.from_numpy(fx["q_mdl"]))
    assert (O.quantise(t64(rs["x_openai_f32"]) * 0.5 + 0.5) != torch.from_numpy(fx["q_openai"])).sum().item() == 0
    assert (O.quantise(t64(rs["x01_mdl_f32"])) != torch.from_numpy(fx["q_mdl"])).sum().item() == 0


def test_oracle_matches_reference_source_mdl_plain_and_full_iwae_loss():
    fx, rs = golden("plain_m5_latent"), refsrc("plain_m5_latent")
    params = torch.from_numpy(fx["params"]).double().requires_grad_(True)
    x = O.normalize_u8(torch.from_numpy(fx["x_u8"]), torch.float32).double()
    lp = O.mdl_plain_log_prob(params, x)
    ll = lp.sum((-1, -2))
    (ll * t64(fx["g_image"])).sum().backward()
    assert (lp.detach() - t64(rs["lp_f64"])).abs().max().item() <= CAST_ABS * 15
    assert_ll_close(ll, rs["ll_f64"], rtol=1e-8)
    assert relnorm(params.grad, t64(rs["grad_f64"])) <= 1e-7
    xs, _ = O.mdl_plain_sample(torch.from_numpy(fx["params"]), torch.from_numpy(fx["u_mix"]), torch.from_numpy(fx["u_log"]))
    xm, _ = O.mdl_plain_sample(torch.from_numpy(fx["params"]), torch.from_numpy(fx["u_mix"]), None)
    assert (xs - t64(rs["x_sample_f64"])).abs().max().item() <= 1e-12
    assert (xm - t64(rs["x_mean_f64"])).abs().max().item() <= 1e-12
    # the class built with its own (low, high, levels) (utils/mdl_plain.py:18)
    pb = torch.from_numpy(fx["params"]).double().requires_grad_(True)
    lpb = O.mdl_plain_log_prob(pb, x, *PLAIN_BINS)
    llb = lpb.sum((-1, -2))
    (llb * t64(fx["g_image"])).sum().backward()
    assert (lpb.detach() - t64(rs["bins_lp_f64"])).abs().max().item() <= CAST_ABS * 15
    assert_ll_close(llb, rs["bins_ll_f64"], rtol=1e-8)
    assert relnorm(pb.grad, t64(rs["bins_grad_f64"])) <= 1e-7
    assert (lpb.detach() - lp.detach()).abs().max().item() > 0.5          # (it is a different density)
    xsb, _ = O.mdl_plain_sample(torch.from_numpy(fx["params"]), torch.from_numpy(fx["u_mix"]), torch.from_numpy(fx["u_log"]),
                                PLAIN_BINS[0], PLAIN_BINS[1])
    assert (xsb - t64(rs["bins_x_sample_f64"])).abs().max().item() <= 1e-12
    assert (xm - t64(rs["bins_x_mean_f64"])).abs().max().item() <= 1e-12  # mean() clips to [-1, 1] whatever low / high (:115)
    # iwae_loss with Normal latents, beta = 0.7 (models/loss.py:26-55 executed by the reference's code)
    import torch.distributions as td
    z = t64(fx["z"]).requires_grad_(True)
    ql, qs = t64(fx["q_loc"]).requires_grad_(True), t64(fx["q_scale"]).requires_grad_(True)
    p2 = torch.from_numpy(fx["params"]).double().requires_grad_(True)
    lpz = td.Normal(0.0, 1.0).log_prob(z).sum(-1)
    lqzx = td.Normal(ql, qs).log_prob(z).sum(-1)
    loss, met = O.iwae_loss(O.modl_log_prob(p2, x), lpz, lqzx, x.shape, beta=0.7)
    loss.backward()
    assert abs(loss.item() - float(rs["full_loss_f64"])) <= 1e-8 * abs(loss.item())
    assert relnorm(p2.grad, t64(rs["full_dparams_f64"])) <= 1e-6
    assert relnorm(z.grad, t64(rs["full_dz_f64"])) <= 1e-9
    assert relnorm(ql.grad, t64(rs["full_dq_loc_f64"])) <= 1e-9 and relnorm(qs.grad, t64(rs["full_dq_scale_f64"])) <= 1e-9
    assert relnorm(lpz, t64(rs["full_lpz_f64"])) <= 1e-12 and relnorm(lqzx, t64(rs["full_lqzx_f64"])) <= 1e-12


# ------------------------------------------------------------------------------------------------------------------
# GPU: the product (classes -> ctypes -> C ABI -> sm_100a kernels) against the reference's code
# ------------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def V(built_lib):
    import vae_mdl_b200
    return vae_mdl_b200


@pytest.mark.gpu
@pytest.mark.parametrize("name,M", MODL)
def test_gpu_modl_classes_match_reference_source(V, name, M):
    from vae_mdl_b200 import functional as F
    fx, rs = golden(name), refsrc(name)
    params = torch.from_numpy(fx["params"]).to(DEV)
    x_u8 = torch.from_numpy(fx["x_u8"]).to(DEV)
    x = V.normalize(x_u8)                                                        # utils/data.py:16
    g_image = torch.from_numpy(fx["g_image"]).to(DEV)
    for cls, key in ((V.MixtureDiscretizedLogistic, "mdl"), (V.MixtureDiscretizedLogisticOpenaiIWAE, "iwae_cls")):
        p = params.clone().requires_grad_(True)
        pxz = cls(p)
        lp = pxz.log_prob(x)
        assert list(lp.shape) == list(params.shape[:-1]) + [1]
        assert (lp.detach().cpu().double()[..., 0] - t64(rs[key + "_lp_f64"])).abs().max().item() < 5e-5
        ll = torch.sum(lp, dim=(-1, -2, -3))                                     # models/loss.py:32
        assert_ll_close(ll, rs["mdl_ll_f64"])
        (ll * g_image).sum().backward()
        assert_grad_close(p.grad, rs[key + "_grad_fixed_f64"], M)
    lp3 = V.MixtureDiscretizedLogisticOpenai(params[0]).log_prob(x * 2.0 - 1.0)
    assert (lp3.cpu().double() - t64(rs["openai_cls_lp_s0_f64"])).abs().max().item() < 5e-5
    tot = V.discretized_mix_logistic_loss(x * 2.0 - 1.0, params[0], sum_all=True)
    assert abs(tot.item() - float(rs["openai_sum_all_s0_f64"])) <= LL_RTOL * abs(float(rs["openai_sum_all_s0_f64"]))
    # the fused loss + gradient step against models/loss.py::iwae_loss run by the reference's code
    loss, lpxz, dp = V.modl_iwae_step(params, x_u8, torch.from_numpy(fx["extra"]).to(DEV))
    assert abs(loss.item() - float(rs["iwae_loss_f64"])) <= LL_RTOL * abs(float(rs["iwae_loss_f64"]))
    assert_ll_close(lpxz, rs["iwae_lpxz_f64"])
    assert_grad_close(dp, rs["iwae_grad_f64"], M)
    lw = t64(rs["iwae_lpxz_f64"]) + t64(fx["extra"])
    lme = V.logmeanexp(lw.to(DEV), axis=0)
    assert (lme.cpu().double() - t64(rs["logmeanexp_axis0_f64"])).abs().max().item() <= LL_RTOL * lw.abs().max().item()


@pytest.mark.gpu
def test_gpu_dl_and_model06_loss_match_reference_source(V):
    import torch.distributions as td
    fx, rs = golden("dl_small"), refsrc("dl_small")
    both = torch.from_numpy(fx["both"]).to(DEV)
    x = V.normalize(torch.from_numpy(fx["x_u8"]).to(DEV))
    loc, ls = both[..., :3].clone().requires_grad_(True), both[..., 3:].clone().requires_grad_(True)
    d = V.DiscretizedLogistic(loc, ls, low=0.0, high=1.0, levels=256.0)          # models/model03.py:97
    lp = d.log_prob(x)
    assert (lp.detach().cpu().double() - t64(rs["lp_f64"])).abs().max().item() < 5e-5
    ll = lp.sum((-1, -2, -3))
    assert_ll_close(ll, rs["ll_f64"])
    (ll * torch.from_numpy(fx["g_image"]).to(DEV)).sum().backward()
    assert relnorm(loc.grad, t64(rs["dloc_f64"])) <= GRAD_RTOL and relnorm(ls.grad, t64(rs["dls_f64"])) <= GRAD_RTOL
    # model06's loss_fn through the package's loss_fn (fused route: Normal latents + DiscretizedLogistic)
    dev = lambda k: torch.from_numpy(rs[k]).to(DEV)
    z1, z2 = dev("m6_z1"), dev("m6_z2")
    S, B = z1.shape[:2]
    loc2, ls2 = both[..., :3].clone().requires_grad_(True), both[..., 3:].clone().requires_grad_(True)
    DT = V.DistributionTuple
    pz = td.Normal(torch.zeros(S, B, 4, device=DEV), torch.ones(S, B, 4, device=DEV))
    pz.axes = [-1]
    qz1x = DT(td.Normal(dev("m6_q1_loc"), dev("m6_q1_scale")), z1, (-1,))
    qz2z1 = DT(td.Normal(dev("m6_q2_loc"), dev("m6_q2_scale")), z2, (-1,))
    pz1z2 = DT(td.Normal(dev("m6_p1_loc"), dev("m6_p1_scale")), None, (-1,))
    pxz1 = DT(V.DiscretizedLogistic(loc2, ls2, low=0.0, high=1.0, levels=256.0), None, (-1, -2, -3))
    loss, met = V.loss_fn(x, pz, qz1x, qz2z1, pz1z2, pxz1)
    loss.backward()
    ref = float(rs["m6_loss_f64"])
    assert abs(loss.item() - ref) <= LL_RTOL * abs(ref)
    assert abs(met["bpd"].item() - float(rs["m6_bpd_f64"])) <= LL_RTOL * abs(float(rs["m6_bpd_f64"]))
    assert relnorm(loc2.grad, t64(rs["m6_dloc_f64"])) <= GRAD_RTOL and relnorm(ls2.grad, t64(rs["m6_dls_f64"])) <= GRAD_RTOL
    for k in ("lpxz", "lqz1x", "lqz2z1", "lpz2", "lpz1z2", "kl1", "kl2"):
        assert relnorm(met[k], t64(rs["m6_%s_f64" % k])) <= LL_RTOL, k


@pytest.mark.gpu
def test_gpu_samplers_match_reference_source(V):
    fx, rs = golden("sample_m10"), refsrc("sample_m10")
    l = torch.from_numpy(fx["l"]).to(DEV)
    u_mix, u_log, u_all = (torch.from_numpy(fx[k]).to(DEV) for k in ("u_mix", "u_log", "u_log_all"))
    q = lambda a01: O.quantise(t64(a01))
    x = V.sample_from_discretized_mix_logistic(l, 10, u_mix=u_mix, u_log=u_log)              # [-1,1]
    assert (x.cpu().double() - t64(rs["x_openai_f64"])).abs().max().item() <= 1e-6
    assert torch.equal(O.quantise(x.cpu().double() * 0.5 + 0.5), q(rs["x_openai_f64"] * 0.5 + 0.5))
    x1 = V.MixtureDiscretizedLogisticOpenai(l).sample(1, u_mix=u_mix[None], u_log=u_log[None])
    assert (x1.cpu().double() - t64(rs["x_openai_cls_f64"])).abs().max().item() <= 1e-6
    x2 = V.MixtureDiscretizedLogisticOpenaiIWAE(l).sample(u_mix=u_mix[None], u_log=u_log[None])
    assert (x2.cpu().double() - t64(rs["x01_iwae_cls_f64"])).abs().max().item() <= 1e-6
    x3, q3, idx3 = V.MixtureDiscretizedLogistic(l).sample(u_mix=u_mix[None], u_log=u_all[None], return_index=True,
                                                          return_quantised=True)
    assert (x3.cpu().double() - t64(rs["x01_mdl_f64"])).abs().max().item() <= 1e-6
    assert torch.equal(q3.cpu(), q(rs["x01_mdl_f64"]))
    assert torch.equal(idx3.cpu().to(torch.uint8), torch.from_numpy(fx["idx"]))
    xd = V.DiscretizedLogistic(l[..., :3].contiguous(), l[..., 3:6].contiguous(), low=0.0, high=1.0, levels=256.0).sample(u=u_log[None])
    assert (xd.cpu().double().reshape(t64(rs["x_dl_f64"]).shape) - t64(rs["x_dl_f64"])).abs().max().item() <= 1e-6


@pytest.mark.gpu
def test_gpu_mdl_plain_and_full_iwae_loss_match_reference_source(V):
    import torch.distributions as td
    fx, rs = golden("plain_m5_latent"), refsrc("plain_m5_latent")
    params = torch.from_numpy(fx["params"]).to(DEV)
    x = V.normalize(torch.from_numpy(fx["x_u8"]).to(DEV))
    p = params.clone().requires_grad_(True)
    d = V.PixelMixtureDiscretizedLogistic(p)
    lp = d.log_prob(x)
    assert (lp.detach().cpu().double() - t64(rs["lp_f64"])).abs().max().item() < 5e-5
    ll = lp.sum((-1, -2))
    assert_ll_close(ll, rs["ll_f64"])
    (ll * torch.from_numpy(fx["g_image"]).to(DEV)).sum().backward()
    assert_grad_close(p.grad, rs["grad_f64"], 5)
    u_mix, u_log = torch.from_numpy(fx["u_mix"]).to(DEV), torch.from_numpy(fx["u_log"]).to(DEV)
    d0 = V.PixelMixtureDiscretizedLogistic(params)
    xs = d0.sample(u_mix=u_mix[None], u_log=u_log[None])
    assert (xs.cpu().double() - t64(rs["x_sample_f64"])).abs().max().item() <= 1e-6
    assert torch.equal(O.quantise(xs.cpu().double()), O.quantise(t64(rs["x_sample_f64"])))
    xm = d0.mean(u_mix=u_mix[None])
    assert (xm.cpu().double() - t64(rs["x_mean_f64"])).abs().max().item() <= 1e-6
    # the class built with its own (low, high, levels) (utils/mdl_plain.py:18): edges at low / high, 64 levels
    pb = params.clone().requires_grad_(True)
    db = V.PixelMixtureDiscretizedLogistic(pb, low=PLAIN_BINS[0], high=PLAIN_BINS[1], levels=PLAIN_BINS[2])
    lpb = db.log_prob(x)
    assert (lpb.detach().cpu().double() - t64(rs["bins_lp_f64"])).abs().max().item() < 5e-5
    llb = db.log_likelihood(x, dtype=torch.float64)
    assert_ll_close(llb, rs["bins_ll_f64"])
    (lpb.sum((-1, -2)) * torch.from_numpy(fx["g_image"]).to(DEV)).sum().backward()
    assert_grad_close(pb.grad, rs["bins_grad_f64"], 5)
    db0 = V.PixelMixtureDiscretizedLogistic(params, low=PLAIN_BINS[0], high=PLAIN_BINS[1], levels=PLAIN_BINS[2])
    xsb = db0.sample(u_mix=u_mix[None], u_log=u_log[None])
    assert (xsb.cpu().double() - t64(rs["bins_x_sample_f64"])).abs().max().item() <= 1e-6
    assert (db0.mean(u_mix=u_mix[None]).cpu().double() - t64(rs["bins_x_mean_f64"])).abs().max().item() <= 1e-6
    # models/loss.py::iwae_loss with Normal latents, beta = 0.7: the package's iwae_loss (fused route) vs the reference's
    z = torch.from_numpy(fx["z"]).to(DEV).requires_grad_(True)
    ql = torch.from_numpy(fx["q_loc"]).to(DEV).requires_grad_(True)
    qs = torch.from_numpy(fx["q_scale"]).to(DEV).requires_grad_(True)
    p2 = params.clone().requires_grad_(True)
    pz = td.Normal(torch.zeros_like(z), torch.ones_like(z))
    pz.axes = [-1]
    qzx = td.Normal(ql, qs)
    qzx.axes = [-1]
    loss, met = V.iwae_loss(x, z, pz, qzx, V.MixtureDiscretizedLogistic(p2), beta=0.7)
    loss.backward()
    ref = float(rs["full_loss_f64"])
    assert abs(loss.item() - ref) <= LL_RTOL * abs(ref)
    assert abs(met["bpd"].item() - float(rs["full_bpd_f64"])) <= LL_RTOL * abs(float(rs["full_bpd_f64"]))
    assert_grad_close(p2.grad, rs["full_dparams_f64"], 5)
    assert relnorm(z.grad, t64(rs["full_dz_f64"])) <= GRAD_RTOL
    assert relnorm(ql.grad, t64(rs["full_dq_loc_f64"])) <= GRAD_RTOL and relnorm(qs.grad, t64(rs["full_dq_scale_f64"])) <= GRAD_RTOL
    for k in ("lpxz", "lqzx", "lpz", "kl"):
        assert relnorm(met[k], t64(rs["full_%s_f64" % k])) <= LL_RTOL, k
