"""Sampler throughput vs n_mix (openai variant, 10,000 32x32 images), GPU only."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vae_mdl_b200 import functional as F, _abi
dev = torch.device("cuda:0")
PEAK = 6549.1
for M in [int(m) for m in os.environ.get("MS", "1,2,4,5,8,10,12,16,20,30,32,64").split(",")]:
    N = max(64, int(10000 * 10 / M)) if M > 10 else 10000
    l = torch.randn(N, 32, 32, 10 * M, device=dev)
    um = torch.rand(N, 32, 32, M, device=dev) * (1 - 2e-5) + 1e-5
    ul = torch.rand(N, 32, 32, 3, device=dev) * (1 - 2e-5) + 1e-5
    def f(): F.modl_sample(l, um, ul, _abi.SAMPLE_OPENAI, _abi.RANGE_SYM, want_quantised=True, want_index=True)
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): f()
    e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 10 * 1e-3
    gb = N * 1024 * (44 * M + 12 + 16) / 1e9
    print(f"M={M:3d} N={N}: {t*1e3:7.3f} ms  {N/t/1e6:6.2f} M img/s  {gb/t:7.1f} GB/s  {gb/t/PEAK*100:5.1f}%", flush=True)
    del l, um, ul
