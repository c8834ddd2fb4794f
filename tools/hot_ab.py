"""Interleaved A/B in ONE hot process (power-capped clocks after the first seconds): two-pass gradient on tensor memory vs two-pass in
shared memory vs one-pass from the forward pass's sums, whole step, GPU only.   python tools/hot_ab.py [workload ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import ModlStep, WORKLOADS
dev = torch.device("cuda:0")
names = sys.argv[1:] or ["cfg5_64_m10", "cfg5_64_m20"]
MODES = {"tensor memory, two-pass": {"VAEMDL_TM": "1", "VAEMDL_STATS": "none"},
         "shared memory, two-pass": {"VAEMDL_TM": "0", "VAEMDL_STATS": "none"},
         "one-pass (forward sums)": {"VAEMDL_TM": "1", "VAEMDL_STATS": "all"}}
def timeit(fn, iters, warm=4):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3
os.environ["VAEMDL_FUSED"] = "0"
for name in names:
    _, S, B, H, W, M = WORKLOADS[name]
    nbuf = max(2, -(-3 * 126 * 2**20 // (S * B * H * W * 40 * M)))
    st = ModlStep(S, B, H, W, M, dev, 7, B, n_buffers=nbuf)
    iters = max(50, int(0.5e6 / (S * B * H * W * M * 2e-5)))   # ~0.5 s per measurement
    for _ in range(3): timeit(st.step, iters)                    # heat up: ~1.5 s of back-to-back steps
    res = {k: [] for k in MODES}
    for rep in range(5):
        for k, env in MODES.items():
            os.environ.update(env)
            res[k].append(timeit(st.step, iters))
    f = lambda v: f"median {sorted(v)[len(v)//2]:8.1f} us  (" + " ".join(f"{x:.0f}" for x in v) + ")"
    print(f"{name} hot ({iters} steps per measurement):", flush=True)
    for k in MODES: print(f"   {k:26s} {f(res[k])}", flush=True)
    del st
