"""Randomised agreement between the oracle (oracle/ref.py) and the reference's OWN modules executed over oracle/tf_shim.

Complements the fixed fixtures (refsrc_*.npz): N random small cases -- shapes, n_mix, parameter scales, observations with
both edge values, (low, high, levels) of the plain class, log-mean-exp axes -- in float64, values and autograd gradients.
Needs /root/reference (run by tests/test_reference_golden.py in a subprocess when it is mounted).

    python tests/golden/fuzz_reference.py [n_cases] [seed]
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_reference_golden import ROOT, import_reference  # noqa: E402

sys.path.insert(0, ROOT)
import oracle as O  # noqa: E402


def main(n_cases=40, seed=0):
    tf, tfp, U, L, M6 = import_reference()
    g = torch.Generator().manual_seed(seed)
    ri = lambda lo, hi: int(torch.randint(lo, hi + 1, (1,), generator=g))
    worst = {"mdl": 0.0, "openai": 0.0, "grad": 0.0, "dl": 0.0, "lme": 0.0, "plain": 0.0}
    for case in range(n_cases):
        S, B, H, W, M = ri(1, 3), ri(1, 3), ri(1, 5), ri(1, 5), ri(1, 6)
        scale = [0.3, 1.0, 3.0][ri(0, 2)]
        params = (torch.randn(S, B, H, W, 10 * M, generator=g) * scale).double()
        params[..., M:] -= 1.5 * (ri(0, 1))                     # sometimes narrow scales / shifted means
        x_u8 = torch.randint(0, 256, (B, H, W, 3), generator=g)
        x_u8.view(-1)[::3] = torch.tensor([0, 255, 7])[torch.randint(0, 3, ((x_u8.numel() + 2) // 3,), generator=g)]
        x = (x_u8.float() / 255.0).double()
        w = torch.randn(S, B, H, W, generator=g).double()
        # utils/mdl.py
        p = params.clone().requires_grad_(True)
        lp_ref = U.MixtureDiscretizedLogistic(tf.Tensor(p)).log_prob(tf.Tensor(x)).t[..., 0]
        (lp_ref * w).sum().backward()
        q = params.clone().requires_grad_(True)
        lp_o = O.modl_log_prob(q, x)[..., 0]
        (lp_o * w).sum().backward()
        worst["mdl"] = max(worst["mdl"], float((lp_ref.detach() - lp_o.detach()).abs().max()) / (3 * M))
        worst["grad"] = max(worst["grad"], float((p.grad - q.grad).norm() / q.grad.norm().clamp_min(1e-300)))
        # utils/mdl_openai_iwae.py
        lp2 = U.MixtureDiscretizedLogisticOpenaiIWAE(tf.Tensor(params)).log_prob(tf.Tensor(x)).t[..., 0]
        worst["openai"] = max(worst["openai"], float((lp2 - O.modl_openai_iwae_log_prob(params, x)[..., 0]).abs().max()))
        # utils/mdl_plain.py
        lp3 = U.PixelMixtureDiscretizedLogistic(tf.Tensor(params)).log_prob(tf.Tensor(x)).t
        worst["plain"] = max(worst["plain"], float((lp3 - O.mdl_plain_log_prob(params, x)).abs().max()) / (3 * M))
        # utils/discretized_logistic.py with random (low, high, levels)
        low, high = [(-1.0, 1.0), (0.0, 1.0), (-0.5, 2.0)][ri(0, 2)]
        levels = float([256, 16, 2][ri(0, 2)])
        loc = torch.rand(S, B, H, W, 3, generator=g).double() * (high - low) + low
        ls = torch.randn(S, B, H, W, 3, generator=g).double() - 1.0
        grid = torch.randint(0, int(levels), (B, H, W, 3), generator=g).double() / (levels - 1.0) * (high - low) + low
        d_ref = U.DiscretizedLogistic(tf.Tensor(loc), tf.Tensor(ls), low=low, high=high, levels=levels).log_prob(tf.Tensor(grid)).t
        worst["dl"] = max(worst["dl"], float((d_ref - O.dlogistic_log_prob(grid, loc, ls, low, high, levels)).abs().max()))
        # utils/utils.py::logmeanexp
        lw = torch.randn(ri(1, 6), ri(1, 4), ri(1, 3), generator=g).double() * 5 - 2e4
        # (the reference subtracts the un-broadcast maximum, utils/utils.py:10-11: only the leading axis works there)
        worst["lme"] = max(worst["lme"], float((U.logmeanexp(tf.Tensor(lw), axis=0).t - O.logmeanexp(lw, 0)).abs().max()))
    # mdl.py / discretized_logistic.py / mdl_plain.py round log(interval_width) to float32 in the low-probability branch
    # (utils/mdl.py:163): 4e-7 per such sub-pixel; everything else is round-off
    limits = {"mdl": 4e-7, "plain": 4e-7, "dl": 4e-7, "openai": 1e-9, "grad": 1e-6, "lme": 1e-9}
    ok = all(worst[k] <= limits[k] for k in worst)
    print(("reference source and oracle agree on %d random cases: " % n_cases if ok else "MISMATCH: ")
          + ", ".join("%s %.2e" % kv for kv in sorted(worst.items())))
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main(int(sys.argv[1]) if len(sys.argv) > 1 else 40, int(sys.argv[2]) if len(sys.argv) > 2 else 0))
