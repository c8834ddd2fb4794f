"""Sampler at config-3 size a few times -- a short target for ncu (GPU only)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vae_mdl_b200 as V
dev = torch.device("cuda:0")
N, M = int(sys.argv[1]) if len(sys.argv) > 1 else 4000, 10
g = torch.Generator(device=dev).manual_seed(0)
l = torch.randn(N, 32, 32, 10 * M, device=dev, generator=g)
um = torch.rand(N, 32, 32, M, device=dev, generator=g) * (1 - 2e-5) + 1e-5
ul = torch.rand(N, 32, 32, 3, device=dev, generator=g) * (1 - 2e-5) + 1e-5
for _ in range(4):
    x, xq, idx = V.sample_from_discretized_mix_logistic(l, M, um, ul, return_quantised=True, return_index=True)
torch.cuda.synchronize()
print(int(idx.sum()))
