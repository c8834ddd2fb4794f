"""Data-parallel sharding of the observation-model path over the GPUs of one box (one process per GPU).

The path shards without any data-path collective: every (importance-sample, image) pair is independent
until the log-mean-exp over samples, and test images are independent of each other.

* training shapes: split the BATCH across ranks -- every rank holds all S samples of its images, so the log-mean-exp
  and its gradient are rank-local; the only exchange is one scalar all-reduce of the partial ELBO sums.
* training shapes with fewer images than ranks (or to balance): split the IMPORTANCE SAMPLES instead -- every rank holds
  S / world samples of ALL images; the log-mean-exp over s then needs one exchange of a (max, sum-exp) pair per image
  (``all_gather`` of ``2 * B`` float64), after which value, ELBO and the softmax weights of the local samples are
  rank-local again (``sample_sharded_iwae_step``).
* 5000-sample evaluation (models/model05.py:168-176): images round-robin over ranks, S streamed in chunks into a
  per-image float64 buffer, ONE all-gather of the per-image results at the very end -- no per-image host sync.

``torch.distributed`` (NCCL on GPUs, gloo in the CPU tests) is only the plumbing for those few floats.
"""
from __future__ import annotations

import os
from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist

__all__ = ["init_from_env", "shard_bounds", "round_robin", "gather_round_robin", "allreduce_sum", "IwaeEvaluator",
           "sharded_modl_iwae_step", "combine_lme_over_ranks", "sample_sharded_iwae_step", "split_sample_tail"]


def init_from_env(backend: Optional[str] = None) -> Tuple[int, int, int]:
    """Reads RANK / WORLD_SIZE / LOCAL_RANK (torchrun); initialises the process group when WORLD_SIZE > 1."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            kw["device_id"] = torch.device("cuda", local_rank)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, world, local_rank


def shard_bounds(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced split of ``range(n)``: the first ``n % world`` ranks get one extra element."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def round_robin(n: int, rank: int, world: int) -> range:
    """Image ``i`` belongs to rank ``i % world``."""
    return range(rank, n, world)


def gather_round_robin(local: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """Inverse of ``round_robin``: every rank passes its ``[len(round_robin(n_total, rank, world))]`` results and gets
    the full ``[n_total]`` vector in image order.  One all-gather of ``ceil(n_total/world)`` elements per rank."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    per = (n_total + world - 1) // world
    padded = torch.zeros(per, dtype=local.dtype, device=local.device)
    padded[: local.numel()] = local
    out = torch.empty(world * per, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded, group=group)
    # out[r*per + j] is image j*world + r
    return out.reshape(world, per).t().reshape(-1)[:n_total].contiguous()


def allreduce_sum(t: torch.Tensor, group=None) -> torch.Tensor:
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


class IwaeEvaluator:
    """Sharded version of ``Model05.test(n_samples)`` (models/model05.py:168-176).

    ``ll_chunk_fn(image_index, s_lo, s_hi, out)`` must write the per-sample log-weights ``log_w[s_lo:s_hi]`` of image
    ``image_index`` into ``out`` (a float64 view of length ``s_hi - s_lo`` on ``device``) -- typically
    ``functional.modl_log_likelihood(decoder(z[s_lo:s_hi]), x[image_index], dtype=float64) + lpz - lqzx``.
    ``lme_fn(log_w [S, n]) -> [n]`` is the log-mean-exp over axis 0 (``functional.logmeanexp`` on the GPU).
    Nothing is synchronised with the host until ``run`` returns.
    """

    def __init__(self, n_samples: int, s_chunk: int, rank: int = 0, world: int = 1, device="cuda", group=None):
        self.S, self.s_chunk, self.rank, self.world = int(n_samples), int(s_chunk), rank, world
        self.device, self.group = device, group

    def run(self, n_images: int, ll_chunk_fn: Callable, lme_fn: Callable):
        mine = list(round_robin(n_images, self.rank, self.world))
        # [n_local, S]: each image's S log-weights are contiguous, so a chunk of samples is a contiguous slice
        log_w = torch.empty((max(1, len(mine)), self.S), dtype=torch.float64, device=self.device)
        for j, i in enumerate(mine):
            row = log_w[j]
            for s_lo in range(0, self.S, self.s_chunk):
                s_hi = min(self.S, s_lo + self.s_chunk)
                ll_chunk_fn(i, s_lo, s_hi, row[s_lo:s_hi])
        if mine:
            llh_local = lme_fn(log_w[: len(mine)].t().contiguous())
        else:
            llh_local = torch.empty(0, device=self.device)
        llh = gather_round_robin(llh_local.float(), n_images, self.group)
        return llh.mean(), llh


def sharded_modl_iwae_step(step_fn: Callable, params_shard, x_shard, extra_shard, b_total: int, group=None,
                           exchange=None):
    """One IWAE observation-model step with the batch split across ranks.  ``step_fn`` is
    ``vae_mdl_b200.modl_iwae_step``; each rank gets its additive share of the loss, the sum over the ranks is the global
    loss.  Gradients stay rank-local (they belong to the rank's own decoder activations).

    ``exchange``: a ``vae_mdl_b200.peer.ElboExchange`` -- the kernel that forms the share then stores it straight into
    every rank's exchange buffer (NVLink P2P) and the global loss is a rank-order sum of those words: no collective
    launch.  Without it: one scalar ``all_reduce`` (NCCL on GPUs, gloo in the CPU tests)."""
    if exchange is not None:
        seq = exchange.attach()
        loss, lpxz, dparams = step_fn(params_shard, x_shard, extra_shard, True, b_total)
        return -exchange.read(seq).clone(), lpxz, dparams
    loss, lpxz, dparams = step_fn(params_shard, x_shard, extra_shard, True, b_total)
    loss = allreduce_sum(loss.clone(), group)
    return loss, lpxz, dparams


def combine_lme_over_ranks(log_w_local: torch.Tensor, s_total: int, group=None):
    """Log-mean-exp over an importance-sample axis that is split across ranks (utils/utils.py:9-11 with the samples of
    axis 0 spread over the group).  ``log_w_local [S_local, B]`` (float64 recommended).  One ``all_gather`` of ``[2, B]``
    per rank: the local maximum and the local ``sum_s exp(log_w - max)``.  Returns ``(lme [B], weights [S_local, B])``
    with ``weights = softmax over ALL ``s_total`` samples, restricted to the local ones`` (what the gradient of the
    log-mean-exp needs).  Every rank gets bit-identical ``lme`` (the ranks' pairs are combined in rank order)."""
    mx = log_w_local.amax(0)
    sm = torch.exp(log_w_local - mx).sum(0)
    pair = torch.stack([mx, sm])                                                  # [2, B]
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        world = dist.get_world_size(group)
        flat = torch.empty(world * pair.numel(), dtype=pair.dtype, device=pair.device)
        dist.all_gather_into_tensor(flat, pair.reshape(-1).contiguous(), group=group)
        allp = flat.reshape((world,) + tuple(pair.shape))
    else:
        allp = pair[None]
    gmax = allp[:, 0].amax(0)
    gsum = (allp[:, 1] * torch.exp(allp[:, 0] - gmax)).sum(0)                     # rank order: identical everywhere
    lme = gmax + torch.log(gsum / float(s_total))                                 # utils/utils.py:11
    weights = torch.exp(log_w_local - gmax) / gsum
    return lme, weights


def sample_sharded_iwae_step(ll_fn: Callable, bwd_fn: Callable, params_shard, x, extra_shard, s_total: int, group=None):
    """One IWAE observation-model step with the IMPORTANCE SAMPLES split across ranks (second row): this rank
    holds ``params_shard [S_local, B, H, W, 10M]`` -- its samples of every image -- and ``extra_shard [S_local, B]``.

    ``ll_fn(params, x) -> [S_local, B]`` float64 per-image log-likelihoods (``functional.modl_log_likelihood(...,
    dtype=torch.float64)``), ``bwd_fn(params, x, g_image) -> dparams`` (``functional.modl_backward``).
    Returns ``(loss = -elbo, lpxz [S_local, B], dparams_shard)``; the loss is the GLOBAL loss on every rank (no further
    collective), the gradient is the rank's own slice of the global gradient."""
    lpxz = ll_fn(params_shard, x)
    if lpxz.is_cuda and lpxz.dtype == torch.float64:
        # device route: two small kernels around the one all_gather (vaemdl_iwae_split_local / _combine)
        elbo, g_ll = split_sample_tail(lpxz, extra_shard, s_total, group)
        dparams = bwd_fn(params_shard, x, g_ll)
        return -elbo, lpxz, dparams
    # host-logic route (CPU tensors over gloo in the tests): the same arithmetic with torch ops
    log_w = lpxz.double() + (extra_shard.double() if extra_shard is not None else 0.0)  # models/loss.py:34
    lme, weights = combine_lme_over_ranks(log_w, s_total, group)
    B = log_w.shape[1]
    elbo = lme.mean()                                                             # models/loss.py:37
    g_ll = (-weights / B).float()                                                 # d(-elbo) / d lpxz
    dparams = bwd_fn(params_shard, x, g_ll.contiguous())
    return (-elbo).reshape(1).float(), lpxz, dparams


def split_sample_tail(lpxz64: torch.Tensor, extra_shard: Optional[torch.Tensor], s_total: int, group=None,
                      pairs_out: Optional[torch.Tensor] = None):
    """The IWAE tail for importance samples split across ranks, on the device: ``lpxz64 [S_local, B]`` float64 CUDA.
    ``vaemdl_iwae_split_local`` -> ONE ``all_gather`` of the ``[2, B]`` float64 (max, sum-exp) pairs ->
    ``vaemdl_iwae_split_combine``.  Returns ``(elbo [1], g_ll [S_local, B])`` float32; ``elbo`` is the global value on
    every rank, ``g_ll = d(-elbo)/d lpxz`` for the local samples."""
    from . import _abi
    from ._abi import check, lib, ptr, stream_ptr
    S_local, B = lpxz64.shape
    dev = lpxz64.device
    ex = _abi.dense_f32(extra_shard, "extra") if extra_shard is not None else None
    ll = lpxz64.contiguous()
    world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
    pair = torch.empty((2, B), dtype=torch.float64, device=dev)
    with _abi.on_device(dev):
        check(lib().vaemdl_iwae_split_local(ptr(ll), ptr(ex), S_local, B, ptr(pair), stream_ptr(dev)), "vaemdl_iwae_split_local")
    if world > 1:
        pairs = pairs_out if pairs_out is not None else torch.empty((world, 2, B), dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(pairs.reshape(-1), pair.reshape(-1), group=group)
    else:
        pairs = pair
    elbo = torch.empty(1, dtype=torch.float32, device=dev)
    g_ll = torch.empty((S_local, B), dtype=torch.float32, device=dev)
    with _abi.on_device(dev):
        check(lib().vaemdl_iwae_split_combine(ptr(ll), ptr(ex), S_local, B, ptr(pairs), world, int(s_total), 0, None, None,
                                              ptr(elbo), ptr(g_ll), stream_ptr(dev)), "vaemdl_iwae_split_combine")
    return elbo, g_ll
