"""Throughput of the plain discretized-logistic sampler (models/model06.py:166 draws x on every forward), GPU only."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vae_mdl_b200 as V
dev = torch.device("cuda:0")
for (S, B, H, W) in [(5, 128, 32, 32), (16, 256, 64, 64)]:
    both = torch.randn(S, B, H, W, 6, device=dev); both[..., :3].uniform_()
    u = torch.rand(S, B, H, W, 3, device=dev) * (1 - 2e-5) + 1e-5
    d = V.DiscretizedLogistic(both[..., :3], both[..., 3:], low=0.0, high=1.0, levels=256.0)
    f = lambda: d.sample(u=u[None])
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): f()
    e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 20 * 1e-3
    n = S * B * H * W * 3
    print(f"{S}x{B}x{H}x{W}: {t*1e6:8.1f} us  {n/t/1e9:6.2f} G elements/s  {n*16/t/1e9:7.1f} GB/s (16 B per element)", flush=True)
