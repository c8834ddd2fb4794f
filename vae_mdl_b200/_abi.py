"""ctypes binding of ``libvaemdl_b200.so`` (the C ABI declared in ``include/vaemdl.h``).

PyTorch tensors are used only as device buffers: every call passes ``tensor.data_ptr()``,
sizes and the current CUDA stream handle.  There is NO fallback: if the shared library has
not been built (``python -c "import __graft_entry__ as g; g.build()"`` or
``make -C vae_mdl_b200/csrc``) importing this module's ``lib()`` raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_longlong, c_size_t, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# (VAEMDL_LIB_PATH: A/B measurements of two builds inside one process launch; never set in normal use)
LIB_PATH = os.environ.get("VAEMDL_LIB_PATH") or os.path.join(_HERE, "libvaemdl_b200.so")

# enums (include/vaemdl.h)
X_F32, X_U8 = 0, 1
RANGE_UNIT, RANGE_SYM = 0, 1
EDGE_MDL, EDGE_OPENAI = 0, 1
SAMPLE_OPENAI, SAMPLE_MDL, SAMPLE_PLAIN = 0, 1, 2
MAX_MIX = 64
MAX_LATENT_TERMS = 4


class LatentTerm(ctypes.Structure):
    """``vaemdl_latent_term`` of include/vaemdl.h."""
    _fields_ = [("z", c_void_p), ("loc", c_void_p), ("scale", c_void_p), ("D", c_int), ("params_per_sample", c_int),
                ("weight", c_float)]



class VaemdlPeer(ctypes.Structure):
    """``VaemdlPeer`` of include/vaemdl.h (the peer-memory exchange of the ELBO shares)."""
    _fields_ = [("slots", ctypes.c_uint64 * 8), ("n_ranks", c_int), ("rank", c_int), ("ring", c_int), ("seq", ctypes.c_uint)]


_LIB = None

# name -> (restype, argtypes); kept in one table so tests can check it against the header
PROTOTYPES = {
    "vaemdl_version": (c_int, []),
    "vaemdl_strerror": (c_char_p, [c_int]),
    "vaemdl_modl_workspace_bytes": (c_size_t, [c_longlong, c_int, c_int]),
    "vaemdl_modl_fwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_longlong, c_int, c_int, c_int, c_int,
                                c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "vaemdl_modl_iwae_fwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_longlong, c_longlong, c_int, c_int,
                                     c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                     c_void_p, c_size_t, c_void_p]),
    "vaemdl_modl_iwae_step": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_longlong, c_longlong, c_int, c_int,
                                      c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                      c_void_p, c_void_p, c_size_t, c_void_p, c_void_p]),
    "vaemdl_modl_plain_iwae_step": (c_int, [c_void_p, c_void_p, c_int, c_int, c_longlong, c_longlong, c_int, c_int,
                                            c_int, c_int, c_float, c_float, c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                            c_void_p, c_void_p, c_void_p, c_size_t, c_void_p, c_void_p]),
    "vaemdl_modl_step_workspace_bytes": (c_size_t, [c_longlong, c_int, c_int]),
    "vaemdl_modl_iwae_fwd_stats": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_longlong, c_longlong, c_int,
                                           c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                           c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "vaemdl_modl_bwd_stats": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_longlong, c_int, c_int, c_int, c_int,
                                      c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "vaemdl_modl_bwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_longlong, c_int, c_int, c_int, c_int,
                                c_void_p, c_void_p, c_void_p, c_void_p]),
    "vaemdl_modl_fwd_bf16": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_longlong, c_int, c_int, c_int, c_int,
                                     c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "vaemdl_modl_iwae_fwd_bf16": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_longlong, c_longlong, c_int,
                                          c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                          c_void_p, c_void_p, c_size_t, c_void_p]),
    "vaemdl_modl_bwd_bf16": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_longlong, c_int, c_int, c_int, c_int,
                                     c_void_p, c_void_p, c_void_p, c_void_p]),
    "vaemdl_modl_iwae_fwd_stats_bf16": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_longlong, c_longlong, c_int,
                                                c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                                c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "vaemdl_modl_bwd_stats_bf16": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_longlong, c_int, c_int, c_int, c_int,
                                           c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "vaemdl_modl_plain_fwd": (c_int, [c_void_p, c_void_p, c_int, c_longlong, c_int, c_int, c_int, c_int,
                                      c_float, c_float, c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "vaemdl_modl_plain_iwae_fwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_longlong, c_longlong, c_int, c_int,
                                           c_int, c_int, c_float, c_float, c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                           c_void_p, c_size_t, c_void_p]),
    "vaemdl_modl_plain_bwd": (c_int, [c_void_p, c_void_p, c_int, c_longlong, c_int, c_int, c_int, c_int,
                                      c_float, c_float, c_float, c_void_p, c_void_p, c_void_p, c_void_p]),
    "vaemdl_dlogistic_workspace_bytes": (c_size_t, [c_longlong, c_longlong]),
    "vaemdl_dlogistic_fwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, c_longlong, c_int, c_longlong,
                                     c_float, c_float, c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "vaemdl_dlogistic_iwae_fwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_longlong, c_longlong,
                                          c_int, c_longlong, c_float, c_float, c_float, c_void_p, c_void_p, c_void_p,
                                          c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "vaemdl_dlogistic_iwae_step": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_longlong, c_longlong,
                                           c_int, c_longlong, c_float, c_float, c_float, c_void_p, c_void_p, c_void_p,
                                           c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p,
                                           c_size_t, c_void_p, c_void_p]),
    "vaemdl_dlogistic_bwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, c_longlong, c_int, c_longlong,
                                     c_float, c_float, c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "vaemdl_logmeanexp_fwd": (c_int, [c_void_p, c_int, c_longlong, c_void_p, c_void_p]),
    "vaemdl_logmeanexp_bwd": (c_int, [c_void_p, c_void_p, c_int, c_longlong, c_void_p, c_void_p]),
    "vaemdl_logmeanexp_fwd_f64": (c_int, [c_void_p, c_int, c_longlong, c_void_p, c_void_p]),
    "vaemdl_logmeanexp_bwd_f64": (c_int, [c_void_p, c_void_p, c_int, c_longlong, c_void_p, c_void_p]),
    "vaemdl_latent_terms_fwd": (c_int, [c_void_p, c_int, c_int, c_longlong, c_void_p, c_void_p, c_void_p, c_void_p]),
    "vaemdl_latent_terms_bwd": (c_int, [c_void_p, c_int, c_int, c_longlong, c_void_p, c_void_p, c_void_p, c_void_p,
                                        c_void_p]),
    "vaemdl_iwae_tail": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_longlong, c_longlong, c_void_p, c_void_p, c_void_p,
                                 c_void_p, c_void_p]),
    "vaemdl_iwae_split_local": (c_int, [c_void_p, c_void_p, c_int, c_longlong, c_void_p, c_void_p]),
    "vaemdl_iwae_split_combine": (c_int, [c_void_p, c_void_p, c_int, c_longlong, c_void_p, c_int, c_int, c_longlong, c_void_p,
                                          c_void_p, c_void_p, c_void_p, c_void_p]),
    "vaemdl_peer_next": (c_int, [c_void_p]),
    "vaemdl_peer_elbo_sum": (c_int, [c_void_p, c_int, c_int, ctypes.c_uint, c_void_p, c_void_p]),
    "vaemdl_modl_sample": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_longlong, c_longlong, c_int, c_int, c_int,
                                   c_void_p, c_void_p, c_void_p, c_void_p]),
    "vaemdl_modl_plain_sample": (c_int, [c_void_p, c_void_p, c_void_p, c_float, c_float, c_int, c_longlong, c_longlong, c_int,
                                         c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "vaemdl_dlogistic_sample": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_longlong, c_float, c_float,
                                        c_void_p, c_void_p]),
    "vaemdl_modl_iwae_step_host": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                           c_void_p, c_void_p, c_void_p, c_void_p, c_int]),
    "vaemdl_host_release": (None, []),
}


class VaemdlError(RuntimeError):
    pass


def lib() -> ctypes.CDLL:
    """Load (once) and return the shared library.  Raises if it is missing -- there is no CPU path."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise VaemdlError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(vae_mdl_b200 has no CPU or PyTorch fallback)"
            )
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(handle, name)  # AttributeError => header / library drift
            fn.restype = res
            fn.argtypes = args
        _LIB = handle
    return _LIB


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().vaemdl_strerror(rc).decode()
        raise VaemdlError(f"{what} failed: {msg} (code {rc})")


def ptr(t):
    """Device pointer of a tensor (``None`` -> NULL)."""
    return None if t is None else c_void_p(t.data_ptr())


def stream_ptr(device=None):
    """Raw handle of the current CUDA stream of ``device`` (the fast private accessor when torch has it: this is called
    for every kernel launch and ``torch.cuda.current_stream`` costs several microseconds)."""
    try:
        idx = device.index if isinstance(device, torch.device) and device.index is not None else torch.cuda.current_device()
        return c_void_p(torch._C._cuda_getCurrentRawStream(idx))
    except AttributeError:  # pragma: no cover  (older / newer torch without the private accessor)
        return c_void_p(torch.cuda.current_stream(device).cuda_stream)


class _NoSwitch:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


_NO_SWITCH = _NoSwitch()


def on_device(device):
    """Context that makes ``device`` current for the C-ABI calls inside it -- a no-op object when it already is (the common
    case; ``torch.cuda.device`` costs a few microseconds per launch otherwise)."""
    if device.index is None or device.index == torch.cuda.current_device():
        return _NO_SWITCH
    return torch.cuda.device(device)


def require_cuda(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise VaemdlError(
            f"{name} must be a CUDA tensor: vae_mdl_b200 runs only on the GPU (no CPU fallback); got device {t.device}"
        )


def dense_f32(t: torch.Tensor, name: str) -> torch.Tensor:
    """float32, contiguous, 16-byte aligned view/copy of ``t`` on its CUDA device."""
    require_cuda(t, name)
    if t.dtype != torch.float32:
        t = t.float()
    t = t.contiguous()
    if t.data_ptr() % 16:
        t = t.clone(memory_format=torch.contiguous_format)
    return t


def dense_param(t: torch.Tensor, name: str):
    """MoDL parameter tensor as the kernels read it: ``(tensor, is_bf16)``.  bfloat16 stays bfloat16 (the kernels widen it
    in shared memory); everything else becomes float32.  Contiguous and 16-byte aligned either way."""
    require_cuda(t, name)
    if t.dtype != torch.bfloat16:
        return dense_f32(t, name), False
    t = t.contiguous()
    if t.data_ptr() % 16:
        t = t.clone(memory_format=torch.contiguous_format)
    return t, True
