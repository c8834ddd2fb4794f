"""Golden vectors produced by EXECUTING THE REFERENCE'S OWN SOURCE (nbip/vae-mdl, mounted at /root/reference).

TensorFlow / TFP cannot be installed here, so the reference's unmodified modules are imported with `oracle/tf_shim`
(a torch-CPU stand-in for the ~45 tf / tfd primitives they call, see oracle/tf_shim/README.md) first on `sys.path`.
Every number written below comes out of the reference's code: `utils/mdl.py`, `utils/mdl_openai.py`,
`utils/mdl_openai_iwae.py`, `utils/discretized_logistic.py`, `utils/mdl_plain.py`, `utils/utils.py::logmeanexp`,
`models/loss.py::iwae_loss / elbo_loss`, `models/model06.py::loss_fn`; gradients are autograd over that executed code
(the `tf.GradientTape` of models/model05.py:141-145).

Inputs are the ones already committed in the sibling fixtures (`modl_*.npz`, `dl_small.npz`, `sample_m10.npz`,
`plain_m5_latent.npz`), so `refsrc_<name>.npz` holds outputs only:  `*_f64` = the reference's code run on float64
inputs, `*_f32` = on float32 inputs (the reference's working precision).

    python tests/golden/make_reference_golden.py [--check]     # needs /root/reference; --check compares, writes nothing

Only this script and tests/test_reference_golden.py::test_fixtures_reproduce (skipped when /root/reference is absent)
ever touch /root/reference; nothing that runs on the GPU box does.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REFERENCE = os.environ.get("VAEMDL_REFERENCE", "/root/reference")


def import_reference():
    """Returns (tf, tfd, utils, loss, model06) with the reference's modules loaded from REFERENCE over the shim."""
    shim = os.path.join(ROOT, "oracle", "tf_shim")
    for p in (REFERENCE, shim):
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    sys.dont_write_bytecode = True  # /root/reference is read-only
    import tensorflow as tf
    import tensorflow_probability as tfp
    import utils as ref_utils
    import models.loss as ref_loss
    import models.model06 as ref_model06
    assert os.path.realpath(ref_utils.__file__).startswith(os.path.realpath(REFERENCE)), ref_utils.__file__
    assert "tf_shim" in tf.__file__
    return tf, tfp, ref_utils, ref_loss, ref_model06


class GivenLogProb:
    """Stand-in for a latent distribution whose summed log-prob is a given [S,B] tensor (the modl_* fixtures carry
    `extra = lpz - lqzx` directly, not latents): `log_prob(z)` returns `value[..., None]`, `axes = [-1]`."""

    def __init__(self, tf, value):
        self._v = tf.Tensor(value.unsqueeze(-1))
        self.axes = [-1]

    def log_prob(self, z):
        return self._v


def _np(t):
    t = t.t if hasattr(t, "t") else t
    return t.detach().numpy()


def modl_outputs(R, fx, dtype):
    tf, tfp, U, L, M6 = R
    tag = "_f64" if dtype == torch.float64 else "_f32"
    params = torch.from_numpy(fx["params"]).to(dtype)
    x = (torch.from_numpy(fx["x_u8"]).to(torch.float32) / 255.0).to(dtype)          # utils/data.py:16
    g_image = torch.from_numpy(fx["g_image"]).to(dtype)
    extra = torch.from_numpy(fx["extra"]).to(dtype)
    out = {}
    # utils/mdl.py class, per-pixel log-prob, per-image sums, gradient for a fixed upstream g_image
    p = params.clone().requires_grad_(True)
    pxz = U.MixtureDiscretizedLogistic(tf.Tensor(p))
    lp = pxz.log_prob(tf.Tensor(x))
    ll = tf.reduce_sum(lp, axis=pxz.axes)
    (ll.t * g_image).sum().backward()
    out["mdl_lp" + tag] = _np(lp)[..., 0]
    out["mdl_ll" + tag] = _np(ll)
    out["mdl_grad_fixed" + tag] = p.grad.numpy().copy()
    # utils/mdl_openai_iwae.py class (needs x with its batch dim)
    p = params.clone().requires_grad_(True)
    pxz2 = U.MixtureDiscretizedLogisticOpenaiIWAE(tf.Tensor(p))
    lp2 = pxz2.log_prob(tf.Tensor(x))
    (tf.reduce_sum(lp2, axis=[-1, -2, -3]).t * g_image).sum().backward()
    out["iwae_cls_lp" + tag] = _np(lp2)[..., 0]
    out["iwae_cls_grad_fixed" + tag] = p.grad.numpy().copy()
    # utils/mdl_openai.py class: 4-D logits, x in [-1,1]; sample s = 0 only
    pxz3 = U.MixtureDiscretizedLogisticOpenai(tf.Tensor(params[0]))
    out["openai_cls_lp_s0" + tag] = _np(pxz3.log_prob(tf.Tensor(x * 2.0 - 1.0)))
    from utils.mdl_openai import discretized_mix_logistic_loss
    out["openai_sum_all_s0" + tag] = _np(discretized_mix_logistic_loss(tf.Tensor(x * 2.0 - 1.0), tf.Tensor(params[0])))
    # models/loss.py::iwae_loss end to end (latent terms supplied as given log-probs), gradient w.r.t. parameters
    p = params.clone().requires_grad_(True)
    pxz4 = U.MixtureDiscretizedLogistic(tf.Tensor(p))
    z = tf.Tensor(torch.zeros(extra.shape + (1,), dtype=dtype))
    loss, met = L.iwae_loss(tf.Tensor(x), z, GivenLogProb(tf, extra), GivenLogProb(tf, torch.zeros_like(extra)), pxz4)
    loss.t.backward()
    out["iwae_loss" + tag] = _np(loss)
    out["iwae_grad" + tag] = p.grad.numpy().copy()
    for k in ("iwae_elbo", "bpd", "lpxz", "kl"):
        out["iwae_" + k + tag] = _np(met[k])
    eloss, emet = L.elbo_loss(tf.Tensor(x), z, GivenLogProb(tf, extra), GivenLogProb(tf, torch.zeros_like(extra)), pxz4)
    out["elbo_loss" + tag] = _np(eloss)
    lw = met["lpxz"].t.detach() + extra
    out["logmeanexp_axis0" + tag] = _np(U.logmeanexp(tf.Tensor(lw), axis=0))
    return out


def dl_outputs(R, fx, dtype):
    tf, tfp, U, L, M6 = R
    tag = "_f64" if dtype == torch.float64 else "_f32"
    both = torch.from_numpy(fx["both"]).to(dtype)
    x = (torch.from_numpy(fx["x_u8"]).to(torch.float32) / 255.0).to(dtype)
    g_image = torch.from_numpy(fx["g_image"]).to(dtype)
    loc = both[..., :3].clone().requires_grad_(True)
    ls = both[..., 3:].clone().requires_grad_(True)
    d = U.DiscretizedLogistic(tf.Tensor(loc), tf.Tensor(ls), low=0.0, high=1.0, levels=256.0)   # models/model03.py:97
    lp = d.log_prob(tf.Tensor(x))
    ll = tf.reduce_sum(lp, axis=[-1, -2, -3])
    (ll.t * g_image).sum().backward()
    out = {"lp" + tag: _np(lp), "ll" + tag: _np(ll), "dloc" + tag: loc.grad.numpy().copy(), "dls" + tag: ls.grad.numpy().copy()}
    # models/model06.py::loss_fn with this observation model and Normal latents (two stochastic layers)
    S, B = both.shape[:2]
    g = torch.Generator().manual_seed(606)
    z1 = torch.randn(S, B, 6, generator=g).to(dtype)
    z2 = torch.randn(S, B, 4, generator=g).to(dtype)
    mk = lambda *s: (torch.randn(*s, generator=g).to(dtype), (torch.rand(*s, generator=g) + 0.5).to(dtype))
    q1, q2, p1 = mk(B, 6), mk(S, B, 4), mk(S, B, 6)
    loc2 = both[..., :3].clone().requires_grad_(True)
    ls2 = both[..., 3:].clone().requires_grad_(True)
    d2 = U.DiscretizedLogistic(tf.Tensor(loc2), tf.Tensor(ls2), low=0.0, high=1.0, levels=256.0)
    tfd = tfp.distributions
    pz = tfd.Normal(tf.Tensor(torch.zeros(S, B, 4, dtype=dtype)), tf.Tensor(torch.ones(S, B, 4, dtype=dtype)))
    pz.axes = [-1]
    DT = U.DistributionTuple
    qz1x = DT(tfd.Normal(tf.Tensor(q1[0]), tf.Tensor(q1[1])), tf.Tensor(z1), (-1,))
    qz2z1 = DT(tfd.Normal(tf.Tensor(q2[0]), tf.Tensor(q2[1])), tf.Tensor(z2), (-1,))
    pz1z2 = DT(tfd.Normal(tf.Tensor(p1[0]), tf.Tensor(p1[1])), None, (-1,))
    pxz1 = DT(d2, None, (-1, -2, -3))
    loss, met = M6.loss_fn(tf.Tensor(x), pz, qz1x, qz2z1, pz1z2, pxz1)
    loss.t.backward()
    out.update({"m6_z1": z1.float().numpy(), "m6_z2": z2.float().numpy(),
                "m6_q1_loc": q1[0].float().numpy(), "m6_q1_scale": q1[1].float().numpy(),
                "m6_q2_loc": q2[0].float().numpy(), "m6_q2_scale": q2[1].float().numpy(),
                "m6_p1_loc": p1[0].float().numpy(), "m6_p1_scale": p1[1].float().numpy(),
                "m6_loss" + tag: _np(loss), "m6_dloc" + tag: loc2.grad.numpy().copy(), "m6_dls" + tag: ls2.grad.numpy().copy()})
    for k in ("iwae_elbo", "bpd", "lpxz", "lqz1x", "lqz2z1", "lpz2", "lpz1z2", "kl1", "kl2"):
        out["m6_" + k + tag] = _np(met[k])
    return out


def sample_outputs(R, fx, dtype):
    tf, tfp, U, L, M6 = R
    tag = "_f64" if dtype == torch.float64 else "_f32"
    l = torch.from_numpy(fx["l"]).to(dtype)
    u_mix, u_log, u_log_all = (torch.from_numpy(fx[k]).to(dtype) for k in ("u_mix", "u_log", "u_log_all"))
    M = l.shape[-1] // 10
    out = {}
    from utils.mdl_openai import sample_from_discretized_mix_logistic
    tfp.push_uniforms(u_mix[None], u_log[None])                    # draw order: utils/mdl_openai.py:173, :188
    out["x_openai" + tag] = _np(sample_from_discretized_mix_logistic(tf.Tensor(l), M))
    tfp.push_uniforms(u_mix[None], u_log[None])
    out["x_openai_cls" + tag] = _np(U.MixtureDiscretizedLogisticOpenai(tf.Tensor(l)).sample(1))        # [1,N,H,W,3] in [-1,1]
    tfp.push_uniforms(u_mix[None], u_log[None])
    out["x01_iwae_cls" + tag] = _np(U.MixtureDiscretizedLogisticOpenaiIWAE(tf.Tensor(l)).sample())     # [N,H,W,3] in [0,1]
    tfp.push_uniforms(u_log_all[None], u_mix[None])                # draw order: utils/mdl.py:213, :236-238
    out["x01_mdl" + tag] = _np(U.MixtureDiscretizedLogistic(tf.Tensor(l)).sample())
    # plain DL sampler (utils/discretized_logistic.py:80-85) on the first 6 channels as [..,3] loc / logscale
    tfp.push_uniforms(u_log[None])
    d = U.DiscretizedLogistic(tf.Tensor(l[..., :3]), tf.Tensor(l[..., 3:6]), low=0.0, high=1.0, levels=256.0)
    out["x_dl" + tag] = _np(d.sample())
    assert tfp.pending_uniforms() == 0
    return out


PLAIN_BINS = (-0.95, 0.9, 64.0)   # a non-default (low, high, levels) for PixelMixtureDiscretizedLogistic


def plain_outputs(R, fx, dtype):
    tf, tfp, U, L, M6 = R
    tag = "_f64" if dtype == torch.float64 else "_f32"
    params = torch.from_numpy(fx["params"]).to(dtype)
    x = (torch.from_numpy(fx["x_u8"]).to(torch.float32) / 255.0).to(dtype)
    g_image = torch.from_numpy(fx["g_image"]).to(dtype)
    p = params.clone().requires_grad_(True)
    d = U.PixelMixtureDiscretizedLogistic(tf.Tensor(p))
    lp = d.log_prob(tf.Tensor(x))
    ll = tf.reduce_sum(lp, axis=[-1, -2])
    (ll.t * g_image).sum().backward()
    out = {"lp" + tag: _np(lp), "ll" + tag: _np(ll), "grad" + tag: p.grad.numpy().copy()}
    u_mix, u_log = torch.from_numpy(fx["u_mix"]).to(dtype), torch.from_numpy(fx["u_log"]).to(dtype)
    d = U.PixelMixtureDiscretizedLogistic(tf.Tensor(params))
    tfp.push_uniforms(u_mix[None], u_log[None])                       # draw order: utils/mdl_plain.py:77-78, :86-88
    out["x_sample" + tag] = _np(d.sample())
    tfp.push_uniforms(u_mix[None])
    out["x_mean" + tag] = _np(d.mean())
    # the same class built with its own bin geometry (utils/mdl_plain.py:18): edges at low / high, 64 levels
    pb = params.clone().requires_grad_(True)
    db = U.PixelMixtureDiscretizedLogistic(tf.Tensor(pb), low=PLAIN_BINS[0], high=PLAIN_BINS[1], levels=PLAIN_BINS[2])
    lpb = db.log_prob(tf.Tensor(x))
    llb = tf.reduce_sum(lpb, axis=[-1, -2])
    (llb.t * g_image).sum().backward()
    out.update({"bins_lp" + tag: _np(lpb), "bins_ll" + tag: _np(llb), "bins_grad" + tag: pb.grad.numpy().copy()})
    db = U.PixelMixtureDiscretizedLogistic(tf.Tensor(params), low=PLAIN_BINS[0], high=PLAIN_BINS[1], levels=PLAIN_BINS[2])
    tfp.push_uniforms(u_mix[None], u_log[None])
    out["bins_x_sample" + tag] = _np(db.sample())
    tfp.push_uniforms(u_mix[None])
    out["bins_x_mean" + tag] = _np(db.mean())
    # models/loss.py::iwae_loss with real Normal latents, beta = 0.7, MoDL (utils/mdl.py) observation model
    z = torch.from_numpy(fx["z"]).to(dtype).requires_grad_(True)
    ql = torch.from_numpy(fx["q_loc"]).to(dtype).requires_grad_(True)
    qs = torch.from_numpy(fx["q_scale"]).to(dtype).requires_grad_(True)
    p2 = params.clone().requires_grad_(True)
    tfd = tfp.distributions
    pz = tfd.Normal(tf.Tensor(torch.zeros_like(z)), tf.Tensor(torch.ones_like(z)))
    pz.axes = [-1]
    qzx = tfd.Normal(tf.Tensor(ql), tf.Tensor(qs))
    qzx.axes = [-1]
    pxz = U.MixtureDiscretizedLogistic(tf.Tensor(p2))
    loss, met = L.iwae_loss(tf.Tensor(x), tf.Tensor(z), pz, qzx, pxz, beta=0.7)
    loss.t.backward()
    out.update({"full_loss" + tag: _np(loss), "full_dparams" + tag: p2.grad.numpy().copy(), "full_dz" + tag: z.grad.numpy().copy(),
                "full_dq_loc" + tag: ql.grad.numpy().copy(), "full_dq_scale" + tag: qs.grad.numpy().copy()})
    for k in ("iwae_elbo", "bpd", "lpxz", "lqzx", "lpz", "kl"):
        out["full_" + k + tag] = _np(met[k])
    assert tfp.pending_uniforms() == 0
    return out


JOBS = [
    ("modl_m10_randn", modl_outputs), ("modl_m5_trained", modl_outputs), ("modl_m30_randn", modl_outputs),
    ("modl_m7_ragged", modl_outputs), ("dl_small", dl_outputs), ("sample_m10", sample_outputs),
    ("plain_m5_latent", plain_outputs),
]


def generate(name, fn, R=None):
    R = R or import_reference()
    fx = np.load(os.path.join(HERE, name + ".npz"))
    out = {}
    for dtype in (torch.float64, torch.float32):
        out.update(fn(R, fx, dtype))
    return out


def main(check=False, only=None):
    R = import_reference()
    worst = 0.0
    for name, fn in JOBS:
        if only and name not in only:
            continue
        out = generate(name, fn, R)
        path = os.path.join(HERE, "refsrc_" + name + ".npz")
        if check:
            old = np.load(path)
            for k, v in out.items():
                d = float(np.max(np.abs(np.asarray(v, np.float64) - np.asarray(old[k], np.float64)))) if np.size(v) else 0.0
                worst = max(worst, d)
                assert d == 0.0, (name, k, d)
        else:
            np.savez_compressed(path, **out)
            print("wrote", path, "(%d arrays, %.1f KB)" % (len(out), os.path.getsize(path) / 1e3))
    if check:
        print("refsrc fixtures reproduce bit for bit from", REFERENCE)


if __name__ == "__main__":
    main(check="--check" in sys.argv, only=[a for a in sys.argv[1:] if not a.startswith("--")])
