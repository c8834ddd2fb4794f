"""vae_mdl_b200 -- B200 (sm_100a) kernels for the observation-model hot path of nbip/vae-mdl.

Exports the same names as the reference's ``utils/__init__.py`` (:1-7) for the classes on the hot path, plus the loss
functions of ``models/loss.py``.  Every computation goes through ``libvaemdl_b200.so`` (C ABI in ``include/vaemdl.h``);
there is no CPU or pure-PyTorch fallback -- a missing library or a CPU tensor raises.
"""
from . import _abi
from .discretized_logistic import DiscretizedLogistic
from .loss import dlogistic_iwae_step, elbo_loss, iwae_loss, loss_fn, modl_iwae_step
from .mdl import MixtureDiscretizedLogistic
from .mdl_plain import PixelMixtureDiscretizedLogistic, get_mixture_params
from .mdl_openai import (MixtureDiscretizedLogisticOpenai, discretized_mix_logistic_loss, int_shape,
                         log_prob_from_logits, log_sum_exp, sample_from_discretized_mix_logistic)
from .mdl_openai_iwae import MixtureDiscretizedLogisticOpenaiIWAE
from .utils import Dist, DistributionTuple, fill_canvas, logmeanexp, normalize, sample_grid, write_ppm

__all__ = [
    "DiscretizedLogistic",
    "MixtureDiscretizedLogistic",
    "MixtureDiscretizedLogisticOpenai",
    "MixtureDiscretizedLogisticOpenaiIWAE",
    "PixelMixtureDiscretizedLogistic",
    "get_mixture_params",
    "discretized_mix_logistic_loss",
    "sample_from_discretized_mix_logistic",
    "log_sum_exp",
    "log_prob_from_logits",
    "int_shape",
    "logmeanexp",
    "Dist",
    "DistributionTuple",
    "fill_canvas",
    "normalize",
    "sample_grid",
    "write_ppm",
    "iwae_loss",
    "elbo_loss",
    "loss_fn",
    "modl_iwae_step",
    "dlogistic_iwae_step",
]
