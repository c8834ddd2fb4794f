"""Interleaved A/B of the one-pass (per-pixel sums handed over) vs two-pass gradient inside one process, GPU only."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import ModlStep, WORKLOADS
dev = torch.device("cuda:0")
names = sys.argv[1:] or ["cfg5_64_m30", "cfg5_128_m30", "cfg1_m5", "cfg5_64_m10"]
def timeit(fn, iters=40, warm=4):
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
    res = {"none": [], "all": []}
    for rep in range(5):
        for mode in ("none", "all"):
            os.environ["VAEMDL_STATS"] = mode
            res[mode].append(timeit(st.step))
    f = lambda v: f"median {sorted(v)[len(v)//2]:8.1f} us  (" + " ".join(f"{x:.0f}" for x in v) + ")"
    print(f"{name}: two-pass {f(res['none'])} | one-pass {f(res['all'])}", flush=True)
    del st
