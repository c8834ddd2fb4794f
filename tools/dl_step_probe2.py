"""BASELINE configs[1] (plain discretized logistic, S=5 x B=128, 32x32x3) through vaemdl_dlogistic_iwae_step under the current
environment switches: us per step with 12 rotating input / gradient tensors.  GPU only."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vae_mdl_b200 import _abi
dev = torch.device("cuda:0")
L = _abi.lib()
S, B, H, W = (int(v) for v in (sys.argv[1:5] if len(sys.argv) >= 5 else (5, 128, 32, 32)))
gen = torch.Generator(device=dev).manual_seed(3)
NB = max(2, -(-3 * 126 * 2**20 // (S * B * H * W * 24)))
pool = []
for _ in range(NB):
    t_ = torch.randn(S, B, H, W, 6, device=dev, generator=gen); t_[..., :3].uniform_(generator=gen); pool.append(t_)
dpool = [torch.empty_like(pool[0]) for _ in range(NB)]
x = torch.randint(0, 256, (B, H, W, 3), dtype=torch.uint8, device=dev, generator=gen)
D = H * W * 3
ll64 = torch.empty(S, B, dtype=torch.float64, device=dev); g_ll = torch.empty(S, B, device=dev)
lme = torch.empty(B, device=dev); elbo = torch.empty(1, device=dev)
wsb = L.vaemdl_dlogistic_workspace_bytes(S * B, D); ws = torch.empty(wsb // 8 + 1, dtype=torch.float64, device=dev)
st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
turn = [0]; nl = ctypes.c_int(0)
def step():
    turn[0] += 1
    both, dboth = pool[turn[0] % NB], dpool[turn[0] % NB]
    rc = L.vaemdl_dlogistic_iwae_step(both.data_ptr(), both.data_ptr() + 12, 3, 6, x.data_ptr(), 1, S, B, 0, B, D, 0.0, 1.0, 256.0,
                                      None, None, ll64.data_ptr(), None, lme.data_ptr(), elbo.data_ptr(), g_ll.data_ptr(),
                                      dboth.data_ptr(), dboth.data_ptr() + 12, 6, ws.data_ptr(), wsb, st, ctypes.byref(nl))
    assert rc == 0, rc
def timeit(fn, iters=200, warm=20):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3
ts = [timeit(step) for _ in range(3)]
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    keep = st
    st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    for _ in range(NB): step()
    st = keep
tg = timeit(g.replay, iters=20, warm=3) / NB
tag = " ".join(f"{k}={v}" for k, v in sorted(os.environ.items()) if k.startswith("VAEMDL_"))
print(f"dl step {S}x{B}x{H}x{W} [{tag}] launches {nl.value}: " + " / ".join(f"{t:.1f}" for t in ts) + f" us per step, graph {tg:.1f} us", flush=True)
