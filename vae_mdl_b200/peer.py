"""The step's one collective over peer memory: every rank's kernel stores its ELBO share straight into every rank's
exchange buffer (NVLink P2P stores through CUDA-IPC mappings); no collective launch after the step.

With the batch split across the GPUs of one box each rank's step ends with its additive share of the ELBO
(``sum_b lme_b / B_total``, models/loss.py:37); the loss is the sum of the shares.  ``ElboExchange`` owns the rank's
``[ring, world]`` buffer of 64-bit words, maps the peers' buffers, and hands the C ABI what it needs
(``vaemdl_peer_next`` / ``vaemdl_peer_elbo_sum``, include/vaemdl.h).  One process per GPU; ``torch.distributed`` is used
once, to exchange the 64-byte IPC handles.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import torch
import torch.distributed as dist

from . import _abi

__all__ = ["ElboExchange"]

_IPC_HANDLE_BYTES = 64
_LAZY_PEER_ACCESS = 1  # cudaIpcMemLazyEnablePeerAccess
_rt = None


class _IpcHandle(ctypes.Structure):
    """``cudaIpcMemHandle_t``: 64 opaque bytes, passed to ``cudaIpcOpenMemHandle`` BY VALUE."""
    _fields_ = [("reserved", ctypes.c_char * _IPC_HANDLE_BYTES)]


def _cudart():
    global _rt
    if _rt is None:
        for name in ("libcudart.so.12", "libcudart.so"):
            try:
                _rt = ctypes.CDLL(name)
                break
            except OSError:
                continue
        if _rt is None:
            raise _abi.VaemdlError("libcudart not found")
        _rt.cudaMalloc.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_size_t]
        _rt.cudaFree.argtypes = [ctypes.c_void_p]
        _rt.cudaMemset.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t]
        _rt.cudaIpcGetMemHandle.argtypes = [ctypes.POINTER(_IpcHandle), ctypes.c_void_p]
        _rt.cudaIpcOpenMemHandle.argtypes = [ctypes.POINTER(ctypes.c_void_p), _IpcHandle, ctypes.c_uint]
        _rt.cudaIpcCloseMemHandle.argtypes = [ctypes.c_void_p]
    return _rt


def _ok(rc: int, what: str):
    if rc != 0:
        raise _abi.VaemdlError(f"{what} failed with CUDA error {rc}")


class ElboExchange:
    """``ex = ElboExchange(device, group)`` once; then per step ``ex.attach()`` right before the call that produces the
    rank's ELBO share (``vaemdl_*_iwae_step`` / ``vaemdl_*_iwae_fwd*`` with ``B_total`` = the global batch) and
    ``ex.read()`` for the global ELBO of that step (a ``[1]`` float32 tensor, summed in rank order, identical on every
    rank).  ``read`` only enqueues a one-warp kernel; nothing synchronises with the host."""

    def __init__(self, device, group=None, ring: int = 64):
        self.device = torch.device(device)
        self.group = group
        self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        self.rank = dist.get_rank(group) if self.world > 1 else 0
        if not 1 <= self.world <= 8:
            raise ValueError("one box: 1..8 ranks")
        self.ring = int(ring)
        self.seq = 0
        self._peer_ptrs = []
        rt = _cudart()
        with _abi.on_device(self.device):
            base = ctypes.c_void_p()
            nbytes = self.ring * self.world * 8
            _ok(rt.cudaMalloc(ctypes.byref(base), nbytes), "cudaMalloc")
            _ok(rt.cudaMemset(base, 0, nbytes), "cudaMemset")  # sequence numbers start at 1: zero means "nothing yet"
            torch.cuda.synchronize(self.device)
            self._base = base
            slots = [None] * self.world
            slots[self.rank] = base.value
            if self.world > 1:
                handle = _IpcHandle()
                _ok(rt.cudaIpcGetMemHandle(ctypes.byref(handle), base), "cudaIpcGetMemHandle")
                gathered = [None] * self.world
                dist.all_gather_object(gathered, bytes(bytearray(handle)), group=group)
                for r, raw in enumerate(gathered):
                    if r == self.rank:
                        continue
                    h = _IpcHandle.from_buffer_copy(raw)
                    p = ctypes.c_void_p()
                    _ok(rt.cudaIpcOpenMemHandle(ctypes.byref(p), h, _LAZY_PEER_ACCESS), "cudaIpcOpenMemHandle")
                    self._peer_ptrs.append(p)
                    slots[r] = p.value
                dist.barrier(group=group)  # every mapping exists before anyone stores
        self._slots = slots
        self._struct = None
        self._out = torch.empty(1, dtype=torch.float32, device=self.device)

    def attach(self) -> int:
        """Attaches the exchange to the next ELBO-producing C-ABI call of this host thread; returns the step's sequence
        number (pass it to ``read``)."""
        self.seq += 1
        p = self._struct
        if p is None:
            p = self._struct = _abi.VaemdlPeer()
            for r in range(self.world):
                p.slots[r] = self._slots[r]
            p.n_ranks, p.rank, p.ring = self.world, self.rank, self.ring
            self._struct_ref = ctypes.byref(p)
            self._next = _abi.lib().vaemdl_peer_next
        p.seq = self.seq
        rc = self._next(self._struct_ref)
        if rc:
            _abi.check(rc, "vaemdl_peer_next")
        return self.seq

    def detach(self):
        _abi.check(_abi.lib().vaemdl_peer_next(None), "vaemdl_peer_next")

    def read(self, seq: Optional[int] = None, out: Optional[torch.Tensor] = None, stream=None) -> torch.Tensor:
        """The global ELBO of step ``seq`` (default: the last attached one): enqueues the rank-order sum of the ``world``
        shares on ``stream`` (default: the current stream).  NaN if a share did not arrive within ~2 s or was overrun by
        a step more than ``ring`` steps later."""
        out = self._out if out is None else out
        sp = ctypes.c_void_p(stream.cuda_stream) if stream is not None else _abi.stream_ptr(self.device)
        with _abi.on_device(self.device):
            _abi.check(_abi.lib().vaemdl_peer_elbo_sum(ctypes.c_void_p(self._base.value), self.world, self.ring,
                                                       int(self.seq if seq is None else seq), _abi.ptr(out), sp),
                       "vaemdl_peer_elbo_sum")
        return out

    def close(self):
        rt = _cudart()
        if getattr(self, "_base", None) is None:
            return
        with _abi.on_device(self.device):
            torch.cuda.synchronize(self.device)
            if self.world > 1:
                dist.barrier(group=self.group)  # nobody stores into a buffer that is about to go away
            for p in self._peer_ptrs:
                rt.cudaIpcCloseMemHandle(p)
            rt.cudaFree(self._base)
        self._peer_ptrs, self._base = [], None

    def __del__(self):  # pragma: no cover - best effort
        try:
            if getattr(self, "_base", None) is not None and self.world == 1:
                self.close()
        except Exception:
            pass
