"""Sweep chunk sizes of the host-buffer step (GPU only)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vae_mdl_b200 import _abi
L = _abi.lib()
S, B, H, W, M = 16, 32, 64, 64, 10
g = torch.Generator().manual_seed(0)
params = torch.empty(S, B, H, W, 10 * M).pin_memory(); params.normal_(generator=g)
x = torch.randint(0, 256, (B, H, W, 3), dtype=torch.uint8, generator=g).pin_memory()
extra = torch.randn(S, B, generator=g).pin_memory()
dparams = torch.empty(S, B, H, W, 10 * M).pin_memory()
ll = torch.empty(S, B).pin_memory(); lme = torch.empty(B).pin_memory(); elbo = torch.empty(1).pin_memory()
npx = S * B * H * W
for chunk in [1, 2, 4, 8, 16, 32]:
    def call():
        rc = L.vaemdl_modl_iwae_step_host(params.data_ptr(), x.data_ptr(), extra.data_ptr(), S, B, H, W, M, dparams.data_ptr(), ll.data_ptr(), lme.data_ptr(), elbo.data_ptr(), chunk)
        assert rc == 0
    call(); call()
    t0 = time.perf_counter()
    for _ in range(5): call()
    dt = (time.perf_counter() - t0) / 5
    print(f"chunk_b={chunk:3d} ({chunk*S*H*W*400/2**20:6.1f} MiB/chunk): {dt*1e3:7.2f} ms/step  {npx/dt/1e6:7.1f} Mpx/s  H2D {params.numel()*4/dt/1e9:5.1f} GB/s", flush=True)
# plain copies for reference
d = torch.empty_like(params, device="cuda")
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5): d.copy_(params, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
print(f"plain H2D: {params.numel()*4/dt/1e9:.1f} GB/s")
t0 = time.perf_counter()
for _ in range(5): dparams.copy_(d, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
print(f"plain D2H: {params.numel()*4/dt/1e9:.1f} GB/s")
s1 = torch.cuda.Stream(); s2 = torch.cuda.Stream(); d2 = torch.empty_like(d)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5):
    with torch.cuda.stream(s1): d.copy_(params, non_blocking=True)
    with torch.cuda.stream(s2): dparams.copy_(d2, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
print(f"simultaneous H2D+D2H: {params.numel()*4/dt/1e9:.1f} GB/s each direction")
