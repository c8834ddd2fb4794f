#!/usr/bin/env python
"""bench.py -- the driver-facing benchmark of the observation-model hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path over one batch of synthetic input that is already resident in HBM:
MoDL per-image log-likelihood (forward kernel) -> fused IWAE tail (log-mean-exp over importance samples, ELBO,
softmax weights) -> MoDL parameter gradient (backward kernel); at N > 1 the step also carries the path's one collective
(the scalar all-reduce of the ELBO shares, vae_mdl_b200/dist.py::sharded_modl_iwae_step).  Rank 0 prints ONE JSON line.

Besides the K timed steps the line carries: `value_sustained` (the same step back to back for >= 3 s, clocks recorded),
`configs0` / `configs0_ref_default` (BASELINE configs[0] and the reference's real default shape, each with its own
roofline block), `sample_split` (the step with the importance samples, not the batch, spread over the ranks),
`iwae_eval_5000is` (BASELINE configs[3], 512 images per GPU), `e2e` (host buffers through the C ABI) and `elbo_check`
(the e2e arm's inputs pushed through the device arm: a failed comparison fails the run).

Headline workload (BASELINE.json configs[4], the largest MoDL fwd+bwd configuration, per-GPU shard):
    64x64x3 images, 10 mixtures, 16 importance samples, 32 images per GPU  (= batch 256 over 8 GPUs), weak scaling.
Other BASELINE configs are available through --workload and are summarised under the "also" key.

metric = px-samples/s: one px-sample is one pixel (3 sub-pixels x M mixtures) of one (importance sample, image) pair.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import math
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

WORKLOADS = {
    # name: (kind, S, B_per_gpu, H, W, M)
    "cfg5_64_m10": ("modl", 16, 32, 64, 64, 10),
    "cfg5_64_m30": ("modl", 16, 32, 64, 64, 30),
    "cfg5_128_m10": ("modl", 16, 32, 128, 128, 10),
    "cfg5_128_m30": ("modl", 16, 32, 128, 128, 30),
    "cfg1": ("modl", 5, 64, 32, 32, 10),
    "cfg1_m5": ("modl", 5, 128, 32, 32, 5),
    "cfg5_64_m5": ("modl", 16, 64, 64, 64, 5),
    "cfg5_64_m20": ("modl", 16, 32, 64, 64, 20),
    "cfg1_m20": ("modl", 5, 64, 32, 32, 20),
    "cfg1_m30": ("modl", 5, 64, 32, 32, 30),
}
DEFAULT_WORKLOAD = "cfg5_64_m10"
L2_BYTES = 126 * 1024 * 1024


def measured_hbm_peak():
    """(GB/s, how): MEASURED_PEAKS.json if the driver wrote it, else the profiling recipe's stated fallback."""
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index: int, period_s: float = 0.002):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        self.period = period_s
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            # the first query of each kind can stall in the driver for tens of milliseconds (seen once at N = 2: one sample in an
            # 80 ms region, the launches of that rank held up with it): make it here, before the timed region starts
            try:
                pynvml.nvmlDeviceGetClockInfo(self.h, pynvml.NVML_CLOCK_SM)
                (getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or pynvml.nvmlDeviceGetCurrentClocksThrottleReasons)(self.h)
            except Exception:
                pass
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
            "hw_power_brake": getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80),
        }
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or getattr(
            nv, "nvmlDeviceGetCurrentClocksThrottleReasons")
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = get_reasons(self.h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(self.period)

    def __enter__(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


_FULL_AFFINITY = None


def bind_near_gpu(index: int):
    """Pins this process to the CPUs NVML reports as local to GPU `index` (same NUMA node / PCIe root), so that the
    pinned host buffers of the end-to-end arm are first-touched next to the GPU that DMAs them.  The previous mask is
    kept for the CPU baseline, which uses every core."""
    global _FULL_AFFINITY
    if os.environ.get("VAEMDL_NO_BIND"):
        return None
    try:
        import pynvml
        pynvml.nvmlInit()
        if _FULL_AFFINITY is None:
            _FULL_AFFINITY = os.sched_getaffinity(0)
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(index))
        return len(os.sched_getaffinity(0))
    except Exception:
        return None


def unbind():
    if _FULL_AFFINITY is not None:
        try:
            os.sched_setaffinity(0, _FULL_AFFINITY)
        except Exception:
            pass


def dist_setup(n_gpus):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))
    return rank, world, local_rank


def barrier(world):
    if world > 1:
        import torch.distributed as dist
        dist.barrier()


_GATE = None


def start_gate(world, dev):
    """After the host-side barrier + synchronize: a tiny all-reduce enqueued on the timing stream, so that every rank's timed
    region begins at the same moment ON THE DEVICE.  Without it the ranks leave the host barrier a few hundred microseconds to
    milliseconds apart (Python, NVML), and with an exchange inside the step the early ranks' timed regions then contain the
    wait for the late ones -- 15-40 % of a 20-step region at 4-8 GPUs.  The begin event is recorded right behind the gate."""
    global _GATE
    if world <= 1:
        return
    import torch.distributed as dist
    if _GATE is None or _GATE.device != dev:
        _GATE = torch.zeros(1, device=dev)
    dist.all_reduce(_GATE)  # (the current stream waits for it; the host does not)


def max_over_ranks(value, world, dev):
    if world == 1:
        return value
    import torch.distributed as dist
    t = torch.tensor([value], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


# ------------------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------------------
class ModlStep:
    """Device-resident buffers + raw C-ABI calls for one fwd -> finish -> bwd step (3 launches)."""

    LAUNCHES_PER_STEP = 3  # modl fwd (tile partials), fused finish (per-image sums + IWAE tail + batch mean), modl bwd

    def __init__(self, S, B, H, W, M, dev, seed, b_total, n_buffers=2):
        from vae_mdl_b200 import _abi
        self.L = _abi.lib()
        self.S, self.B, self.H, self.W, self.M, self.dev, self.b_total = S, B, H, W, M, dev, b_total
        gen = torch.Generator(device=dev).manual_seed(seed)
        # Consecutive steps alternate between n_buffers parameter tensors: the forward kernel of step k+1 must not find
        # lines of its input in L2 that step k left there (in training every step scores NEW decoder outputs).
        self.pool = [torch.randn(S, B, H, W, 10 * M, device=dev, generator=gen) for _ in range(n_buffers)]  # utils/mdl.py:295
        self.k = 0
        self.params = self.pool[0]
        self.x = torch.randint(0, 256, (B, H, W, 3), dtype=torch.uint8, device=dev, generator=gen)
        self.extra = torch.randn(S, B, device=dev, generator=gen)                          # beta*(lpz - lqzx)
        self.dparams = torch.empty_like(self.params)
        self.ll64 = torch.empty(S, B, dtype=torch.float64, device=dev)
        self.g_ll = torch.empty(S, B, device=dev)
        self.lme = torch.empty(B, device=dev)
        self.elbo = torch.empty(1, device=dev)
        self.ws_bytes = self.L.vaemdl_modl_step_workspace_bytes(S * B, H, W)
        self.ws = torch.empty(self.ws_bytes // 8 + 1, dtype=torch.float64, device=dev)
        # (mixture sum, logit normaliser) per pixel-sample: written by the forward kernel, read by the backward kernel,
        # which then scales every gradient in the pass that forms it (what vaemdl_modl_iwae_step does internally)
        self.stats = torch.empty(S * B * H * W * 2, dtype=torch.float32, device=dev)
        self.stream = torch.cuda.current_stream(dev)
        self.st = ctypes.c_void_p(self.stream.cuda_stream)
        self.n_px = S * B * H * W

    def fwd(self):
        # forward kernel + fused finish kernel: lpxz (float64), log-mean-exp, elbo, g_ll = d(-elbo)/d lpxz
        if LINK is not None:
            LINK.attach()
        rc = self.L.vaemdl_modl_iwae_fwd_stats(self.params.data_ptr(), self.x.data_ptr(), 1, 0, 0, self.S, self.B,
                                               self.b_total, self.B, self.H, self.W, self.M, self.extra.data_ptr(), None,
                                               self.ll64.data_ptr(), None, self.lme.data_ptr(), self.elbo.data_ptr(),
                                               self.g_ll.data_ptr(), self.stats.data_ptr(), self.ws.data_ptr(),
                                               self.ws_bytes, self.st)
        assert rc == 0, rc

    def bwd(self):
        rc = self.L.vaemdl_modl_bwd_stats(self.params.data_ptr(), self.x.data_ptr(), 1, 0, 0, self.S * self.B, self.B,
                                          self.H, self.W, self.M, self.g_ll.data_ptr(), None, self.stats.data_ptr(),
                                          self.dparams.data_ptr(), self.st)
        assert rc == 0, rc

    def next_input(self):
        self.k += 1
        self.params = self.pool[self.k % len(self.pool)]

    def step(self):
        """One call of vaemdl_modl_iwae_step: a single cooperative launch for small training shapes (BASELINE configs[0]),
        forward + finish + backward (3 launches) otherwise; self.launches holds what the library enqueued."""
        self.next_input()
        n = ctypes.c_int(0)
        if LINK is not None:
            LINK.attach()
        rc = self.L.vaemdl_modl_iwae_step(self.params.data_ptr(), self.x.data_ptr(), 1, 0, 0, self.S, self.B, self.b_total,
                                          self.B, self.H, self.W, self.M, self.extra.data_ptr(), None, self.ll64.data_ptr(),
                                          None, self.lme.data_ptr(), self.elbo.data_ptr(), self.g_ll.data_ptr(),
                                          self.dparams.data_ptr(), self.ws.data_ptr(), self.ws_bytes, self.st, ctypes.byref(n))
        assert rc == 0, rc
        self.launches = n.value


PROBE_EVERY = 4  # every 4th timed step carries the per-kernel CUDA events


class ElboLink:
    """The path's one collective at N > 1 (models/loss.py:37 over a batch split across ranks): every rank's additive share
    of the ELBO -> the global value.  Preferred: the peer-memory exchange (vae_mdl_b200/peer.py) -- the kernel that forms
    the share stores it into every rank's buffer over NVLink, a one-warp kernel adds the shares in rank order; if the IPC
    mappings cannot be set up on this box (any rank), all ranks fall back to a 4-byte NCCL all-reduce per step."""

    def __init__(self, world, dev):
        self.world, self.dev, self.ex, self.kind, self.why = world, dev, None, None, None
        self.pending = None
        self.lag = int(os.environ.get("VAEMDL_BENCH_LINK_LAG", "1"))
        if world <= 1:
            if os.environ.get("VAEMDL_BENCH_LINK1"):  # measurement aid: the exchange machinery with a single rank
                from vae_mdl_b200.peer import ElboExchange
                self.ex = ElboExchange(dev)
                self.kind = "peer memory (single rank, measurement aid)"
                self.global_elbo = torch.zeros(1, device=dev)
                self.world = 2
            return
        import torch.distributed as dist
        ok = 1
        try:
            if os.environ.get("VAEMDL_BENCH_NCCL"):
                raise RuntimeError("VAEMDL_BENCH_NCCL set")
            from vae_mdl_b200.peer import ElboExchange
            self.ex = ElboExchange(dev)
        except Exception as e:  # pragma: no cover - depends on the box
            ok, self.why = 0, repr(e)
        flag = torch.tensor([ok], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            if self.ex is not None:
                self.ex.close()
                self.ex = None
            self.kind = "nccl all_reduce(sum) of the ELBO share, 4 bytes, every step" + (f" (peer memory unavailable: {self.why})" if self.why else "")
        else:
            self.kind = ("peer memory: the kernel that forms the ELBO share stores it into every rank's exchange buffer "
                         "(CUDA-IPC mappings, NVLink P2P) + a one-warp rank-order sum kernel, every step")
        self.global_elbo = torch.zeros(1, device=dev)

    def attach(self):      # right before the call that produces the share
        if self.ex is not None:
            self.seq = self.ex.attach()

    def finish(self, elbo):  # right after the step's launches, same stream
        if self.world <= 1:
            return
        if self.ex is not None:
            # The shares of step k are stored by the step's own kernels (every rank, over NVLink).  Their rank-order sum is
            # enqueued `lag` steps later (default 1): by then every peer's word has long arrived, so no rank ever spins on a
            # slower one inside its critical path.  flush() enqueues what is still outstanding.
            if self.lag <= 0:
                self.ex.read(self.seq, out=self.global_elbo)
            else:
                if self.pending is not None:
                    self.ex.read(self.pending, out=self.global_elbo)
                self.pending = self.seq
        else:
            import torch.distributed as dist
            self.global_elbo.copy_(elbo)
            dist.all_reduce(self.global_elbo, op=dist.ReduceOp.SUM)

    def flush(self):
        if self.ex is not None and self.pending is not None:
            self.ex.read(self.pending, out=self.global_elbo)
            self.pending = None

    def close(self):
        self.flush()
        if self.ex is not None:
            self.ex.close()
            self.ex = None


LINK = None  # set in main()


def allreduce_elbo(step, world):
    if LINK is not None:
        LINK.finish(step.elbo)


def run_sustained(step: ModlStep, seconds, ms_per_step, world, dev, sampler_index):
    """The same step back to back for >= `seconds` (no per-kernel probes, no host sync inside): what the kernels deliver
    once the GPU has warmed up and settled on its sustained clocks.  The step count is fixed up front from the short
    run's time per step, so every rank enqueues the same number of steps."""
    n = max(10, int(math.ceil(seconds * 1e3 / ms_per_step * 1.08)))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(sampler_index, period_s=0.01)
    barrier(world)
    torch.cuda.synchronize(dev)
    with sampler as clk:
        start_gate(world, dev)
        e0.record(step.stream)
        for _ in range(n):
            step.next_input()
            step.fwd()
            step.bwd()
            allreduce_elbo(step, world)
        if LINK is not None:
            LINK.flush()
        e1.record(step.stream)
        torch.cuda.synchronize(dev)
    barrier(world)
    return {"steps": n, "total_ms": e0.elapsed_time(e1), "clocks": clk.summary()}


def run_device_resident(step: ModlStep, steps, warmup, world, dev, sampler_index):
    """Timed region: barrier + sync, K steps, sync + barrier; ONE pair of CUDA events brackets all K steps (`value`).
    Every PROBE_EVERY-th step also records an event between the forward + finish launches and the backward launch --
    the per-kernel durations `roofline` is computed from come out of the same timed region.  (An event record between two
    launches ends the programmatic dependent-launch chain there, which costs that step ~2 %: hence not on every step.)"""
    for _ in range(warmup):
        step.next_input()
        step.fwd()
        step.bwd()
        allreduce_elbo(step, world)
    torch.cuda.synchronize(dev)
    probes = [k for k in range(steps) if k % PROBE_EVERY == 0]
    ev = {k: [torch.cuda.Event(enable_timing=True) for _ in range(3)] for k in probes}
    e_begin, e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(sampler_index)  # (NVML set-up before the barrier: it takes a different time on every rank)
    barrier(world)
    torch.cuda.synchronize(dev)
    with sampler as clk:
        start_gate(world, dev)
        t0 = time.perf_counter()
        e_begin.record(step.stream)
        for k in range(steps):
            step.next_input()
            if k in ev:
                ev[k][0].record(step.stream)
                step.fwd()
                ev[k][1].record(step.stream)
                step.bwd()
                ev[k][2].record(step.stream)
            else:
                step.fwd()
                step.bwd()
            allreduce_elbo(step, world)
        if LINK is not None:
            LINK.flush()
        e_end.record(step.stream)
        torch.cuda.synchronize(dev)
        wall = time.perf_counter() - t0
    barrier(world)
    total_ms = e_begin.elapsed_time(e_end)
    fwd_ms = sum(e[0].elapsed_time(e[1]) for e in ev.values()) / len(ev)
    bwd_ms = sum(e[1].elapsed_time(e[2]) for e in ev.values()) / len(ev)
    return {"total_ms": total_ms, "wall_ms": wall * 1e3, "fwd_ms": fwd_ms, "bwd_ms": bwd_ms, "clocks": clk.summary(),
            "probed_steps": len(ev)}


def run_e2e(S, B, H, W, M, steps, warmup, world, dev, seed):
    """The same step through the C-ABI host-buffer entry point: pinned host params/x/extra in, gradients/ll/lme/elbo
    out, every step; copies are inside the timed region."""
    from vae_mdl_b200 import _abi
    L = _abi.lib()
    g = torch.Generator().manual_seed(seed)
    params = torch.empty(S, B, H, W, 10 * M).pin_memory()
    params.normal_(generator=g)
    x = torch.randint(0, 256, (B, H, W, 3), dtype=torch.uint8, generator=g).pin_memory()
    extra = torch.randn(S, B, generator=g).pin_memory()
    dparams = torch.empty(S, B, H, W, 10 * M).pin_memory()
    ll = torch.empty(S, B).pin_memory()
    lme = torch.empty(B).pin_memory()
    elbo = torch.empty(1).pin_memory()

    def call():
        rc = L.vaemdl_modl_iwae_step_host(params.data_ptr(), x.data_ptr(), extra.data_ptr(), S, B, H, W, M,
                                          dparams.data_ptr(), ll.data_ptr(), lme.data_ptr(), elbo.data_ptr(), 0)
        assert rc == 0, rc

    for _ in range(max(1, min(warmup, 3))):
        call()
    torch.cuda.synchronize(dev)
    barrier(world)
    t0 = time.perf_counter()
    for _ in range(steps):
        call()  # synchronous: returns when the outputs are in host memory
    torch.cuda.synchronize(dev)
    dt = time.perf_counter() - t0
    barrier(world)
    L.vaemdl_host_release()
    h2d = params.numel() * 4 + x.numel() + extra.numel() * 4
    d2h = dparams.numel() * 4 + ll.numel() * 4 + lme.numel() * 4
    check = elbo_check(L, S, B, H, W, M, dev, params, x, extra, dparams, ll, float(elbo.item()))
    return dt, h2d, d2h, check


def elbo_check(L, S, B, H, W, M, dev, params_h, x_h, extra_h, dparams_h, ll_h, elbo_host):
    """The SAME inputs through both arms: the host-buffer arm's pinned tensors are copied to the device and pushed through
    the device-resident entry point (vaemdl_modl_iwae_step, same batch normaliser); ELBO, per-image log-likelihoods and
    the gradient must agree.  Raises (the bench fails) when they do not."""
    p = params_h.to(dev)
    xd = x_h.to(dev)
    ex = extra_h.to(dev)
    dp = torch.empty_like(p)
    ll64 = torch.empty(S, B, dtype=torch.float64, device=dev)
    g_ll, lme, el = torch.empty(S, B, device=dev), torch.empty(B, device=dev), torch.empty(1, device=dev)
    wsb = L.vaemdl_modl_step_workspace_bytes(S * B, H, W)
    ws = torch.empty(wsb // 8 + 1, dtype=torch.float64, device=dev)
    n = ctypes.c_int(0)
    rc = L.vaemdl_modl_iwae_step(p.data_ptr(), xd.data_ptr(), 1, 0, 0, S, B, B, B, H, W, M, ex.data_ptr(), None,
                                 ll64.data_ptr(), None, lme.data_ptr(), el.data_ptr(), g_ll.data_ptr(), dp.data_ptr(),
                                 ws.data_ptr(), wsb, ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream), ctypes.byref(n))
    assert rc == 0, rc
    torch.cuda.synchronize(dev)
    elbo_dev = float(el.item())
    rel_elbo = abs(elbo_dev - elbo_host) / abs(elbo_dev)
    rel_ll = float(((ll_h.to(dev).double() - ll64).abs() / ll64.abs()).max().item())
    # the host arm cuts the batch into chunks and runs the two-pass gradient kernel on each; same arithmetic per pixel
    rel_grad = float(((dparams_h.to(dev) - dp).norm() / dp.norm()).item())
    out = {"device": elbo_dev, "e2e": elbo_host, "rel_elbo": rel_elbo, "rel_ll_max": rel_ll, "rel_grad": rel_grad,
           "tolerance": {"elbo": 1e-5, "ll": 1e-5, "grad": 1e-4}, "same_inputs": True,
           "how": "the e2e arm's pinned host tensors copied to the device and run through vaemdl_modl_iwae_step"}
    out["ok"] = bool(rel_elbo <= 1e-5 and rel_ll <= 1e-5 and rel_grad <= 1e-4)
    if not out["ok"]:
        raise RuntimeError(f"elbo_check failed: {out}")
    return out


def run_eval(world, rank, dev, peak, n_local=512, S=5000, H=32, W=32, M=10):
    """BASELINE configs[3]: test-set IWAE evaluation with 5000 importance samples (models/model05.py:168-176), images
    round-robin over ranks (vae_mdl_b200/dist.py).  Per image: ONE forward launch over the [5000,1,32,32,10M] decoder
    output (2.05 GB at M=10) + ONE finish launch (per-sample sums, log-mean-exp) writing llh[i]; no host sync until the
    single all-gather at the end.  Two parameter buffers alternate (4.1 GB >> L2).  Returns img/s over all ranks."""
    from vae_mdl_b200 import _abi
    from vae_mdl_b200 import dist as vdist
    L = _abi.lib()
    gen = torch.Generator(device=dev).manual_seed(77 + rank)
    pool = [torch.randn(S, 1, H, W, 10 * M, device=dev, generator=gen) for _ in range(2)]
    xs = torch.randint(0, 256, (n_local, H, W, 3), dtype=torch.uint8, device=dev, generator=gen)
    llh = torch.empty(n_local, device=dev)
    ws_bytes = L.vaemdl_modl_workspace_bytes(S, H, W)
    ws = torch.empty(ws_bytes // 8 + 1, dtype=torch.float64, device=dev)
    st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)

    def sweep():
        for i in range(n_local):
            rc = L.vaemdl_modl_iwae_fwd(pool[i & 1].data_ptr(), xs[i].data_ptr(), 1, 0, 0, S, 1, 0, 1, H, W, M, None, None,
                                        None, None, llh[i:].data_ptr(), None, None, ws.data_ptr(), ws_bytes, st)
            assert rc == 0, rc
        return vdist.gather_round_robin(llh, n_local * world)

    for i in range(8):  # warm-up: a few images, not the whole sweep
        rc = L.vaemdl_modl_iwae_fwd(pool[i & 1].data_ptr(), xs[i].data_ptr(), 1, 0, 0, S, 1, 0, 1, H, W, M, None, None,
                                    None, None, llh[i:].data_ptr(), None, None, ws.data_ptr(), ws_bytes, st)
        assert rc == 0, rc
    torch.cuda.synchronize(dev)
    barrier(world)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = sweep()
    e1.record()
    torch.cuda.synchronize(dev)
    barrier(world)
    t = max_over_ranks(e0.elapsed_time(e1), world, dev) * 1e-3
    bpd = float(-out.mean().item() / (math.log(2.0) * H * W * 3))
    n_total = n_local * world
    return {"metric": "5000-IS IWAE test evaluation img/s", "images_per_s": n_total / t, "n_images": n_total,
            "importance_samples": S, "n_mix": M, "ms_per_image_per_gpu": t / n_local * 1e3,
            "frac_of_hbm_peak_per_gpu": n_local * S * H * W * 40 * M / t / 1e9 / peak,
            "full_test_set_26032_images_s": 26032 / (n_total / t), "bpd_of_synthetic_params": bpd,
            "launches_per_image": 2, "collective": "one all_gather of the per-image results at the end"}


def traffic_from_profiles(key, field):
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture (profiles/traffic.json);
    None when no capture of that workload has been committed."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get(key, {}).get(field)
    except Exception:
        return None


def run_small_shape(name, world, dev, peak, steps=60, warm=6):
    """BASELINE configs[0] (`cfg1`: model05 MoDL loss, 10 mixtures, batch 64 x 5 importance samples, 32x32x3) and the
    reference's real default (`cfg1_m5`: n_mix 5, batch 128; models/model05.py:60, utils/data.py:20) through
    vaemdl_modl_iwae_step: ONE cooperative launch per step (forward, IWAE finish, gradient).  Inputs rotate over enough
    buffers that a step's parameters were last touched > 3 x L2 bytes ago.  At N > 1 every rank runs its own batch shard
    and the step carries the ELBO all-reduce.  The one kernel IS the step, so its roofline comes from the step time."""
    _, S, B, H, W, M = WORKLOADS[name]
    nbuf = max(2, -(-3 * L2_BYTES // (S * B * H * W * 40 * M)))
    st = ModlStep(S, B, H, W, M, dev, 7 + int(os.environ.get("RANK", "0")), B * world, n_buffers=nbuf)
    for _ in range(warm):
        st.step()
        allreduce_elbo(st, world)
    torch.cuda.synchronize(dev)
    barrier(world)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start_gate(world, dev)
    e0.record(st.stream)
    for _ in range(steps):
        st.step()
        allreduce_elbo(st, world)
    if LINK is not None:
        LINK.flush()
    e1.record(st.stream)
    torch.cuda.synchronize(dev)
    barrier(world)
    t = max_over_ranks(e0.elapsed_time(e1), world, dev) * 1e-3 / steps
    alg = st.n_px * 120 * M
    out = {"workload": name, "config": {"S": S, "B_per_gpu": B, "H": H, "W": W, "n_mix": M,
                                        "input_rotation": f"{nbuf} parameter tensors ({nbuf * st.n_px * 40 * M / 2**20:.0f} MiB) take turns"},
           "metric": "MoDL fwd+bwd px-samples/s", "value": world * st.n_px / t, "unit": "px-samples/s", "n_gpus": world,
           "steps": steps, "us_per_step": t * 1e6, "launches_per_step": st.launches,
           "roofline": {"bound": "hbm", "kernel": "modl_step_kernel (one cooperative launch: forward + IWAE finish + gradient)"
                        if st.launches == 1 else "forward + finish + gradient launches",
                        "achieved": alg / t / 1e9, "peak": peak, "unit": "GB/s", "frac": alg / t / 1e9 / peak,
                        "traffic": traffic_from_profiles(name, "modl_step_dram_bytes_per_launch"),
                        "algorithmic_bytes_per_launch": alg}}
    if world == 1:
        def on_stream(sp):
            keep, st.st = st.st, sp
            for _ in range(len(st.pool)):  # one graph = one pass over every buffer of the rotation
                st.step()
            st.st = keep
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            on_stream(ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize(dev)
        e0.record()
        for _ in range(20):
            g.replay()
        e1.record()
        torch.cuda.synchronize(dev)
        tg = e0.elapsed_time(e1) * 1e-3 / 20 / len(st.pool)
        out["cuda_graph_us_per_step"] = tg * 1e6
        out["cuda_graph_frac_of_hbm_peak"] = alg / tg / 1e9 / peak
    del st
    torch.cuda.empty_cache()
    return out


def run_sample_split(S_total, B_per_gpu, H, W, M, steps, warm, world, rank, dev, peak):
    """The step with the IMPORTANCE SAMPLES, not the batch, spread over the ranks (SURVEY 8e second row;
    vae_mdl_b200/dist.py::sample_sharded_iwae_step): every rank holds S_total / world samples of ALL B_per_gpu * world
    images -- the same pixel-samples per GPU as the batch split -- and the log-mean-exp over s needs one exchange per
    step: an all_gather of the (max, sum exp) pair of every image (16 * B bytes per rank), inside the timed step.
    5 launches: forward, per-image sums, local pairs, [all_gather], combine (lme, ELBO, local softmax weights), gradient."""
    import torch.distributed as dist
    from vae_mdl_b200 import _abi
    if S_total % world:
        return {"skipped": f"{S_total} importance samples do not split evenly over {world} ranks"}
    L = _abi.lib()
    S, B = S_total // world, B_per_gpu * world
    gen = torch.Generator(device=dev).manual_seed(4321 + rank)
    genx = torch.Generator(device=dev).manual_seed(99)            # every rank scores the same images
    pool = [torch.randn(S, B, H, W, 10 * M, device=dev, generator=gen) for _ in range(2)]
    x = torch.randint(0, 256, (B, H, W, 3), dtype=torch.uint8, device=dev, generator=genx)
    extra = torch.randn(S, B, device=dev, generator=gen)
    dparams = torch.empty_like(pool[0])
    ll64 = torch.empty(S, B, dtype=torch.float64, device=dev)
    pair = torch.empty(2, B, dtype=torch.float64, device=dev)
    pairs = torch.empty(world, 2, B, dtype=torch.float64, device=dev) if world > 1 else pair
    g_ll, lme, elbo = torch.empty(S, B, device=dev), torch.empty(B, device=dev), torch.empty(1, device=dev)
    wsb = L.vaemdl_modl_workspace_bytes(S * B, H, W)
    ws = torch.empty(wsb // 8 + 1, dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream(dev)
    st = ctypes.c_void_p(stream.cuda_stream)
    k = [0]

    def step():
        k[0] += 1
        p = pool[k[0] & 1]
        rc = L.vaemdl_modl_fwd(p.data_ptr(), x.data_ptr(), 1, 0, 0, S * B, B, H, W, M, None, None, ll64.data_ptr(),
                               ws.data_ptr(), wsb, st)
        rc |= L.vaemdl_iwae_split_local(ll64.data_ptr(), extra.data_ptr(), S, B, pair.data_ptr(), st)
        assert rc == 0, rc
        if world > 1:
            dist.all_gather_into_tensor(pairs.view(-1), pair.view(-1))
        rc = L.vaemdl_iwae_split_combine(ll64.data_ptr(), extra.data_ptr(), S, B, pairs.data_ptr(), world, S_total, 0, None,
                                         lme.data_ptr(), elbo.data_ptr(), g_ll.data_ptr(), st)
        rc |= L.vaemdl_modl_bwd(p.data_ptr(), x.data_ptr(), 1, 0, 0, S * B, B, H, W, M, g_ll.data_ptr(), None,
                                dparams.data_ptr(), st)
        assert rc == 0, rc

    for _ in range(warm):
        step()
    torch.cuda.synchronize(dev)
    barrier(world)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start_gate(world, dev)
    e0.record(stream)
    for _ in range(steps):
        step()
    e1.record(stream)
    torch.cuda.synchronize(dev)
    barrier(world)
    t = max_over_ranks(e0.elapsed_time(e1), world, dev) * 1e-3 / steps
    n_px = S * B * H * W
    elbo_all = float(elbo.item())   # the GLOBAL value on every rank
    return {"metric": "MoDL fwd+bwd px-samples/s, importance samples split over the ranks", "value": world * n_px / t,
            "unit": "px-samples/s", "n_gpus": world, "steps": steps, "ms_per_step": t * 1e3, "scaling": "weak",
            "config": {"S_total": S_total, "S_per_gpu": S, "B_all_ranks": B, "H": H, "W": W, "n_mix": M},
            "launches_per_step": 5, "frac_of_hbm_peak_per_gpu": n_px * 120 * M / t / 1e9 / peak,
            "collective": f"one all_gather of [2,{B}] float64 (max, sum-exp) pairs per step, inside the timed step"
                          if world > 1 else "none at N=1 (the pairs stay local)",
            "elbo_global": elbo_all}


def also_workloads(dev, peak):
    """Short, untimed-by-the-driver summaries of the other BASELINE configs (rank 0, N=1 only)."""
    import vae_mdl_b200 as V
    from vae_mdl_b200 import functional as F
    out = {}

    def timeit(fn, iters, warm=3):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1) / iters * 1e-3

    def time_as_graph(fn_with_stream, iters):
        """The same launches captured once into a CUDA graph and replayed (the C-ABI calls only enqueue on the stream
        they are given, so they capture as they are): removes the host's per-launch cost from launch-bound shapes."""
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn_with_stream(ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
        return timeit(g.replay, iters)

    for name in ["cfg5_64_m5", "cfg5_64_m20", "cfg5_64_m30", "cfg5_128_m10", "cfg5_128_m30"]:
        _, S, B, H, W, M = WORKLOADS[name]
        # small shapes rotate over enough buffers that a step's input was last touched > 2 x L2 bytes ago
        nbuf = max(2, -(-3 * L2_BYTES // (S * B * H * W * 40 * M)))
        st = ModlStep(S, B, H, W, M, dev, 7, B, n_buffers=nbuf)
        t = timeit(st.step, 20)
        out[name] = {"px_samples_per_s": st.n_px / t, "us_per_step": t * 1e6, "launches_per_step": st.launches,
                     "algorithmic_GBs": st.n_px * 120 * M / t / 1e9, "frac_of_hbm_peak": st.n_px * 120 * M / t / 1e9 / peak}
        del st
        torch.cuda.empty_cache()
    # bfloat16 parameters / gradient at the headline shape: same launches, half the parameter bytes.
    # NOT the headline (the reference computes in float32); reported to show what the narrower interface costs / buys.
    try:
        _, S, B, H, W, M = WORKLOADS[DEFAULT_WORKLOAD]
        L = ModlStep(1, 1, 8, 8, M, dev, 1, 1).L
        gen = torch.Generator(device=dev).manual_seed(11)
        pool = [torch.randn(S, B, H, W, 10 * M, device=dev, generator=gen).bfloat16() for _ in range(3)]
        xb = torch.randint(0, 256, (B, H, W, 3), dtype=torch.uint8, device=dev, generator=gen)
        dpb = torch.empty_like(pool[0])
        ll64 = torch.empty(S, B, dtype=torch.float64, device=dev)
        gl, lme, el = torch.empty(S, B, device=dev), torch.empty(B, device=dev), torch.empty(1, device=dev)
        wsb = L.vaemdl_modl_workspace_bytes(S * B, H, W)
        ws = torch.empty(wsb // 8 + 1, dtype=torch.float64, device=dev)
        stats = torch.empty(S * B * H * W * 2, device=dev)
        sp = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        k = [0]

        def bf16_step():
            k[0] += 1
            p = pool[k[0] % 3]
            rc = L.vaemdl_modl_iwae_fwd_stats_bf16(p.data_ptr(), xb.data_ptr(), 1, 0, 0, S, B, B, B, H, W, M, None, None,
                                                   ll64.data_ptr(), None, lme.data_ptr(), el.data_ptr(), gl.data_ptr(),
                                                   stats.data_ptr(), ws.data_ptr(), wsb, sp)
            assert rc == 0, rc
            rc = L.vaemdl_modl_bwd_stats_bf16(p.data_ptr(), xb.data_ptr(), 1, 0, 0, S * B, B, H, W, M, gl.data_ptr(), None,
                                              stats.data_ptr(), dpb.data_ptr(), sp)
            assert rc == 0, rc
        t = timeit(bf16_step, 20)
        n_px = S * B * H * W
        out[DEFAULT_WORKLOAD + "_bf16_params"] = {
            "px_samples_per_s": n_px / t, "us_per_step": t * 1e6, "launches_per_step": 3,
            "algorithmic_GBs": n_px * 60 * M / t / 1e9, "frac_of_hbm_peak": n_px * 60 * M / t / 1e9 / peak,
            "note": "NOT the headline (the reference computes in float32): bf16 parameters in, bf16 gradient out (60*M bytes "
                    "per px-sample), float32 arithmetic.  The tile stays bfloat16 in shared memory (two slots per warp), "
                    "component pairs are widened as they are read, the gradient is formed in one pass from the forward "
                    "pass's per-pixel sums and rounded once (vaemdl_modl_iwae_fwd_stats_bf16 / vaemdl_modl_bwd_stats_bf16)"}
        del pool, dpb
    except Exception as e:  # pragma: no cover
        out[DEFAULT_WORKLOAD + "_bf16_params"] = {"error": repr(e)}
    # config 2: plain discretized logistic fwd + IWAE tail + bwd, S=5 x B=128, 32x32x3 (models/model03.py shapes),
    # raw C-ABI calls on preallocated buffers; the un-split [..,6] conv output is read in place (ld = 6)
    S, B, H, W = 5, 128, 32, 32
    gen = torch.Generator(device=dev).manual_seed(3)
    # 12 input tensors (15.7 MB each, 189 MB > L2) and 12 gradient tensors take turns: no call finds its input in L2
    pool = []
    for _ in range(12):
        t_ = torch.randn(S, B, H, W, 6, device=dev, generator=gen)
        t_[..., :3].uniform_(generator=gen)
        pool.append(t_)
    dpool = [torch.empty_like(pool[0]) for _ in range(12)]
    x = torch.randint(0, 256, (B, H, W, 3), dtype=torch.uint8, device=dev, generator=gen)
    from vae_mdl_b200 import _abi
    L = _abi.lib()
    D = H * W * 3
    ll64 = torch.empty(S, B, dtype=torch.float64, device=dev)
    g_ll = torch.empty(S, B, device=dev)
    lme = torch.empty(B, device=dev)
    elbo = torch.empty(1, device=dev)
    wsb = L.vaemdl_dlogistic_workspace_bytes(S * B, D)
    ws = torch.empty(wsb // 8 + 1, dtype=torch.float64, device=dev)
    st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    turn = [0]
    n_launch = ctypes.c_int(0)

    def ptrs():
        turn[0] += 1
        both, dboth = pool[turn[0] % 12], dpool[turn[0] % 12]
        return both.data_ptr(), both.data_ptr() + 12, dboth.data_ptr(), dboth.data_ptr() + 12

    def dl_three(sp=st):
        p_loc, p_ls, d_loc, d_ls = ptrs()
        rc = L.vaemdl_dlogistic_iwae_fwd(p_loc, p_ls, 3, 6, x.data_ptr(), 1, S, B, 0, B, D, 0.0, 1.0, 256.0, None, None,
                                         ll64.data_ptr(), None, lme.data_ptr(), elbo.data_ptr(), g_ll.data_ptr(),
                                         ws.data_ptr(), wsb, sp)
        rc |= L.vaemdl_dlogistic_bwd(p_loc, p_ls, 3, 6, x.data_ptr(), 1, S * B, B, D, 0.0, 1.0, 256.0, g_ll.data_ptr(),
                                     None, d_loc, d_ls, 6, sp)
        assert rc == 0, rc

    def dl_step(sp=st):
        p_loc, p_ls, d_loc, d_ls = ptrs()
        rc = L.vaemdl_dlogistic_iwae_step(p_loc, p_ls, 3, 6, x.data_ptr(), 1, S, B, 0, B, D, 0.0, 1.0, 256.0, None, None,
                                          ll64.data_ptr(), None, lme.data_ptr(), elbo.data_ptr(), g_ll.data_ptr(),
                                          d_loc, d_ls, 6, ws.data_ptr(), wsb, sp, ctypes.byref(n_launch))
        assert rc == 0, rc

    def graph_of(fn):
        def on_stream(sp):
            for _ in range(12):
                fn(sp)
        return time_as_graph(on_stream, 20) / 12

    n_sub = S * B * H * W
    t1 = timeit(dl_step, 100)          # vaemdl_dlogistic_iwae_step as the library runs it (three launches)
    launches = n_launch.value
    tg1 = graph_of(dl_step)
    prev = os.environ.get("VAEMDL_FUSED")
    os.environ["VAEMDL_FUSED"] = "1"   # the opt-in cooperative one-launch kernel, for comparison
    tc = timeit(dl_step, 100)
    launches_c = n_launch.value
    tgc = graph_of(dl_step)
    if prev is None:
        del os.environ["VAEMDL_FUSED"]
    else:
        os.environ["VAEMDL_FUSED"] = prev
    out["cfg2_dl_step"] = {"us_per_step": t1 * 1e6, "px_samples_per_s": n_sub / t1, "launches_per_step": launches,
                           "algorithmic_GBs": n_sub * 72 / t1 / 1e9, "frac_of_hbm_peak": n_sub * 72 / t1 / 1e9 / peak,
                           "cuda_graph_us_per_step": tg1 * 1e6, "cuda_graph_frac_of_hbm_peak": n_sub * 72 / tg1 / 1e9 / peak,
                           "api": "vaemdl_dlogistic_iwae_step: forward, fused finish, gradient kernel launched programmatically "
                                  "behind the finish kernel (its first tile is loaded and evaluated while the finish kernel "
                                  "runs); 12 input / gradient tensors rotate (189 MB > L2)",
                           "cooperative_one_launch_us_per_step": tc * 1e6, "cooperative_launches": launches_c,
                           "cooperative_cuda_graph_us_per_step": tgc * 1e6,
                           "floor": "an EMPTY cooperative kernel with two grid barriers costs 6.2 us on this GPU, three empty "
                                    "launches 9 us (tools/grid_sync_probe.cu); the 47 MB of traffic are 7.2 us at the HBM roofline"}
    del pool, dpool
    # config 3: sampling from supplied uniforms, the full 10,000 x 32 x 32 images of BASELINE configs[2]
    N, M = 10000, 10
    l = torch.randn(N, 32, 32, 10 * M, device=dev, generator=gen)
    um = torch.rand(N, 32, 32, M, device=dev, generator=gen) * (1 - 2e-5) + 1e-5
    ul = torch.rand(N, 32, 32, 3, device=dev, generator=gen) * (1 - 2e-5) + 1e-5
    t = timeit(lambda: V.sample_from_discretized_mix_logistic(l, M, um, ul, return_quantised=True, return_index=True), 5)
    out["cfg3_sampling_10000img"] = {"images_per_s": N / t, "ms": t * 1e3, "algorithmic_GBs": N * 1024 * 468 / t / 1e9,
                                     "frac_of_hbm_peak": N * 1024 * 468 / t / 1e9 / peak,
                                     "bytes_per_pixel": "400 params + 40 u_mix + 12 u_log read, 12 float + 3 uint8 + 1 index written"}
    del l, um, ul
    return out


def cpu_baseline_sample(S, H, W, M, budget_s=12.0):
    """Times the float32 op-for-op restatement of the reference (oracle/ref.py, incl. autograd backward) on the host
    cores, on a bounded sample of the workload: whole images are added until one pass takes ~1 s."""
    import oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    g = torch.Generator().manual_seed(0)

    def one_pass(Bs):
        params = torch.randn(S, Bs, H, W, 10 * M, generator=g).requires_grad_(True)
        x = O.normalize_u8(torch.randint(0, 256, (Bs, H, W, 3), dtype=torch.uint8, generator=g))
        extra = torch.randn(S, Bs, generator=g)
        t0 = time.perf_counter()
        loss, _ = O.iwae_loss(O.modl_log_prob(params, x), extra, torch.zeros_like(extra), x.shape)
        loss.backward()
        return time.perf_counter() - t0

    Bs = 1
    t = one_pass(Bs)  # warm-up + calibration
    t = one_pass(Bs)
    while t < 0.5 and Bs < 64:
        Bs *= 2
        t = one_pass(Bs)
    reps = max(2, min(20, int(budget_s / max(t, 1e-3))))
    times = [one_pass(Bs) for _ in range(reps)]
    best = sorted(times)[len(times) // 2]
    return {"value": S * Bs * H * W / best, "unit": "px-samples/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{reps} passes of fwd+bwd over {Bs} image(s) x {S} importance samples, {H}x{W}x3, M={M} "
                      f"(median {best * 1e3:.0f} ms/pass; torch-CPU float32 restatement of the TF graph, autograd backward)"}


def workload_config(name, wl):
    """`config` of the JSON line -- the same dict for both arms (ours and --impl reference)."""
    kind, S, B, H, W, M = wl
    n_px = S * B * H * W
    return {"workload": name, "S": S, "B_per_gpu": B, "H": H, "W": W, "n_mix": M,
            "px_samples_per_step_per_gpu": n_px,
            "l2": (f"inputs larger than L2 (params {n_px * 40 * M / 2**20:.0f} MiB + grads {n_px * 40 * M / 2**20:.0f} MiB "
                   "per step vs 126 MiB L2)") if n_px * 40 * M > L2_BYTES else "inputs NOT larger than L2",
            "input_rotation": "2 parameter tensors alternate between steps (no L2 carry-over of a step's input)",
            "step": "modl_fwd (tile partials, float64) -> fused finish (per-image ll, log-mean-exp, elbo, softmax weights) "
                    "-> modl_bwd [-> all-reduce of the ELBO shares at N > 1]; inputs resident in HBM"}


# ------------------------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU implementation of the path (restated; TensorFlow cannot run in this image)
# ------------------------------------------------------------------------------------------------------------------
def run_reference(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle as O
    kind, S, B, H, W, M = wl
    torch.set_num_threads(os.cpu_count() or 1)
    g = torch.Generator().manual_seed(0)
    Bs = 2  # bounded sample per step: 2 images x S importance samples
    params = torch.randn(S, Bs, H, W, 10 * M, generator=g)
    x = O.normalize_u8(torch.randint(0, 256, (Bs, H, W, 3), dtype=torch.uint8, generator=g))
    extra = torch.randn(S, Bs, generator=g)

    def step():
        p = params.clone().requires_grad_(True)
        loss, _ = O.iwae_loss(O.modl_log_prob(p, x), extra, torch.zeros_like(extra), x.shape)
        loss.backward()
        return loss.item()

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    n_px = S * Bs * H * W
    value = n_px * args.steps / dt
    sample = (f"each step = fwd+bwd over {Bs} images x {S} importance samples of the workload "
              f"({H}x{W}x3, M={M}); torch-CPU float32 op-for-op restatement of the reference TF graph "
              f"(oracle/ref.py), autograd backward, all host threads")
    line = {
        "impl": "reference", "metric": "MoDL fwd+bwd px-samples/s", "value": value, "unit": "px-samples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.workload, wl),
        "cpu_baseline": {"value": value, "unit": "px-samples/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": "px-samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "reference_sample_images_per_step": Bs,
        "note": "TensorFlow/TFP are not installable in this image; this is the restated reference (kind=port)",
    }
    emit(line)


_JSON_FD = None


def _guard_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints a version banner when the box
    exports NCCL_DEBUG=VERSION), so everything that reaches file descriptor 1 from here on is sent to stderr and the
    JSON line goes to the descriptor stdout had at start-up."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_JSON_FD, data)


def main():
    _guard_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 = min(steps, 10)")
    ap.add_argument("--no-also", action="store_true", help="skip the summaries of the other BASELINE configs")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-eval", action="store_true", help="skip the 5000-IS evaluation (BASELINE configs[3]) leg")
    ap.add_argument("--no-small", action="store_true", help="skip BASELINE configs[0] / the reference-default shape")
    ap.add_argument("--no-split", action="store_true", help="skip the sample-split variant of the step")
    ap.add_argument("--sustain-s", type=float, default=3.0, help="length of the sustained arm in seconds (0 = skip)")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl)
        return
    if args.warmup < 3:
        args.warmup = 3

    rank, world, local_rank = dist_setup(args.gpus)
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    bound_cpus = bind_near_gpu(local_rank)
    kind, S, B, H, W, M = wl
    peak, peak_how = measured_hbm_peak()

    global LINK
    LINK = ElboLink(world, dev)
    step = ModlStep(S, B, H, W, M, dev, seed=1234 + rank, b_total=B * world)
    res = run_device_resident(step, args.steps, args.warmup, world, dev, local_rank)
    total_s = max_over_ranks(res["total_ms"], world, dev) * 1e-3
    n_px = step.n_px
    value = world * n_px * args.steps / total_s
    bwd_s = max_over_ranks(res["bwd_ms"], world, dev) * 1e-3
    fwd_s = max_over_ranks(res["fwd_ms"], world, dev) * 1e-3

    # the same step back to back for >= 3 s: sustained clocks, warm GPU
    sustained = None
    if args.sustain_s > 0:
        sus = run_sustained(step, args.sustain_s, total_s / args.steps * 1e3, world, dev, local_rank)
        sus_s = max_over_ranks(sus["total_ms"], world, dev) * 1e-3
        sustained = {"value": world * n_px * sus["steps"] / sus_s, "unit": "px-samples/s", "steps": sus["steps"],
                     "seconds": sus_s, "ms_per_step": sus_s / sus["steps"] * 1e3,
                     "frac_of_hbm_peak_per_gpu": n_px * 120 * M * sus["steps"] / sus_s / 1e9 / peak, "clocks": sus["clocks"]}

    # roofline of the dominant kernel (backward: reads the 40M-byte row, writes the 40M-byte gradient row)
    bwd_bytes = n_px * 80 * M
    fwd_bytes = n_px * 40 * M
    one_pass = M == 5 or (M == 30 and n_px >= 1200000)  # (what csrc/modl_launch.cuh::stats_supported chooses)
    bwd_kernel = ("modl_tile_kernel<BWD, ST> (one-pass gradient from the forward pass's per-pixel sums)" if one_pass else
                  "modl_tile_tm_kernel (vaemdl_modl_bwd: two-pass gradient, tile in tensor memory)")
    roofline = {"bound": "hbm", "kernel": bwd_kernel, "achieved": bwd_bytes / bwd_s / 1e9,
                "peak": peak, "peak_source": peak_how, "unit": "GB/s", "frac": bwd_bytes / bwd_s / 1e9 / peak,
                "traffic": traffic_from_profiles(args.workload, "modl_bwd_dram_bytes_per_launch"),
                "traffic_source": "profiles/traffic.json (one ncu --set full capture of this kernel; not re-measured here)",
                "algorithmic_bytes_per_launch": bwd_bytes,
                "fwd_plus_finish": {"achieved": fwd_bytes / fwd_s / 1e9, "frac": fwd_bytes / fwd_s / 1e9 / peak,
                                    "algorithmic_bytes_per_launch": fwd_bytes},
                "step": {"achieved": n_px * 120 * M * args.steps / total_s / 1e9,
                         "frac": n_px * 120 * M * args.steps / total_s / 1e9 / peak,
                         "frac_of_nominal_8TBs": n_px * 120 * M * args.steps / total_s / 1e9 / 8000.0}}
    del step
    torch.cuda.empty_cache()

    # BASELINE configs[0] and the reference's real default shape: first-class, each with its own roofline block
    small = {}
    if not args.no_small:
        for key, name in (("configs0", "cfg1"), ("configs0_ref_default", "cfg1_m5")):
            small[key] = run_small_shape(name, world, dev, peak)

    # the importance samples (not the batch) spread over the ranks: the variant with a real exchange step
    split = None
    if not args.no_split:
        split = run_sample_split(S, B, H, W, M, min(args.steps, 30), max(3, min(args.warmup, 5)), world, rank, dev, peak)
        torch.cuda.empty_cache()

    link_kind = LINK.kind
    LINK.close()
    LINK = None
    # end to end through the host-buffer C-ABI call; then the SAME inputs through the device arm (elbo_check)
    e2e_steps = args.e2e_steps or min(args.steps, 10)
    dt, h2d, d2h, check = run_e2e(S, B, H, W, M, e2e_steps, args.warmup, world, dev, seed=99 + rank)
    dt = max_over_ranks(dt, world, dev)
    e2e = {"value": world * n_px * e2e_steps / dt, "unit": "px-samples/s", "h2d_bytes_per_step": h2d,
           "d2h_bytes_per_step": d2h, "steps": e2e_steps, "ms_per_step": dt / e2e_steps * 1e3,
           "api": "vaemdl_modl_iwae_step_host (pinned host buffers; H2D params+x+extra, D2H grads+ll+lme+elbo each step)",
           "host_cpus_bound_near_gpu": bound_cpus,
           "achieved_GBs_per_direction_per_gpu": h2d / (dt / e2e_steps) / 1e9}
    try:  # the box's measured copy ceiling with this many ranks copying both ways at once (tools/pcie_probe.py)
        with open(os.path.join(ROOT, "profiles", "e2e_pcie.json")) as f:
            ceil = json.load(f)["per_n"].get(str(world), {}).get("both", {}).get("per_gpu_GBs_per_direction")
        if ceil:
            e2e["copy_ceiling_GBs_per_direction_per_gpu"] = ceil
            e2e["frac_of_copy_ceiling"] = e2e["achieved_GBs_per_direction_per_gpu"] / ceil
            e2e["ceiling_source"] = "profiles/e2e_pcie.json (simultaneous H2D + D2H, all ranks; the aggregate saturates at 2 GPUs)"
    except Exception:
        pass
    torch.cuda.empty_cache()

    ev = None
    if not args.no_eval:
        ev = run_eval(world, rank, dev, peak)

    cpu = None
    also = None
    if rank == 0 and world == 1:
        if not args.no_cpu_baseline:
            unbind()
            cpu = cpu_baseline_sample(S, H, W, M)
        if not args.no_also:
            try:
                also = also_workloads(dev, peak)
            except Exception as exc:  # pragma: no cover - the headline line must still be printed
                also = {"error": repr(exc)}

    if rank == 0:
        line = {
            "metric": "MoDL fwd+bwd px-samples/s", "value": value, "unit": "px-samples/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_s / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.workload, wl),
            "clocks": res["clocks"], "e2e": e2e, "gpu_launches": ModlStep.LAUNCHES_PER_STEP * args.steps,
            "roofline": roofline, "cpu_baseline": cpu,
            "collective_in_step": link_kind,
            "kernel_ms": {"fwd_plus_finish": res["fwd_ms"], "bwd": res["bwd_ms"], "probed_steps": res["probed_steps"],
                          "how": "CUDA events around the two launch groups on every 4th step of the timed region"},
            "elbo_check": check,
        }
        if sustained is not None:
            line["value_sustained"] = sustained["value"]
            line["sustained"] = sustained
        line.update(small)
        if split is not None:
            line["sample_split"] = split
        if ev is not None:
            line["iwae_eval_5000is"] = ev
        if also is not None:
            line["also"] = also
        emit(line)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
