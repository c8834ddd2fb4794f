"""BASELINE configs[0] / the reference-default shape through vaemdl_modl_iwae_step under the current environment switches:
us per step (rotating inputs), as separate launches and as the one-launch step.  GPU only.
    python tools/cfg1_probe.py [cfg1|cfg1_m5] [iters]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import ModlStep, WORKLOADS
dev = torch.device("cuda:0")
name = sys.argv[1] if len(sys.argv) > 1 else "cfg1"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 100
_, S, B, H, W, M = WORKLOADS[name]
nbuf = max(2, -(-3 * 126 * 2**20 // (S * B * H * W * 40 * M)))
st = ModlStep(S, B, H, W, M, dev, 1, B, n_buffers=nbuf)
def timeit(fn, iters=iters, warm=10):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3
ts = [timeit(st.step) for _ in range(3)]
tag = " ".join(f"{k}={v}" for k, v in sorted(os.environ.items()) if k.startswith("VAEMDL_"))
print(f"{name} [{tag}] launches {st.launches}: " + " / ".join(f"{t:.1f}" for t in ts) + " us per step", flush=True)
