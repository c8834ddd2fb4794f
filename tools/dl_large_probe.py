"""Plain discretized-logistic forward / backward kernels at a large shape (16 x 256 x 64 x 64 x 3), GPU only."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vae_mdl_b200 import _abi
dev = torch.device("cuda:0")
L = _abi.lib()
PEAK = 6549.1
def timeit(fn, iters=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3
for (S, B, H, W) in [(16, 256, 64, 64), (5, 128, 32, 32)]:
    pool = [torch.randn(S, B, H, W, 6, device=dev) for _ in range(3)]
    for t in pool: t[..., :3].uniform_()
    dboth = torch.empty_like(pool[0])
    x = torch.randint(0, 256, (B, H, W, 3), dtype=torch.uint8, device=dev)
    D = H * W * 3; n = S * B * H * W
    ll64 = torch.empty(S, B, dtype=torch.float64, device=dev); g_ll = torch.randn(S, B, device=dev)
    wsb = L.vaemdl_dlogistic_workspace_bytes(S * B, D)
    ws = torch.empty(wsb // 8 + 1, dtype=torch.float64, device=dev)
    st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    k = [0]
    def f():
        k[0] += 1; b = pool[k[0] % 3]
        rc = L.vaemdl_dlogistic_fwd(b.data_ptr(), b.data_ptr() + 12, 3, 6, x.data_ptr(), 1, S * B, B, D, 0.0, 1.0, 256.0, None, None, ll64.data_ptr(), ws.data_ptr(), wsb, st); assert rc == 0, rc
    def g():
        k[0] += 1; b = pool[k[0] % 3]
        rc = L.vaemdl_dlogistic_bwd(b.data_ptr(), b.data_ptr() + 12, 3, 6, x.data_ptr(), 1, S * B, B, D, 0.0, 1.0, 256.0, g_ll.data_ptr(), None, dboth.data_ptr(), dboth.data_ptr() + 12, 6, st); assert rc == 0, rc
    tf_, tb = timeit(f), timeit(g)
    print(f"{S}x{B}x{H}x{W}: fwd {tf_*1e6:8.1f} us {n*24/tf_/1e9:7.1f} GB/s {n*24/tf_/1e9/PEAK*100:5.1f}% | bwd {tb*1e6:8.1f} us {n*48/tb/1e9:7.1f} GB/s {n*48/tb/1e9/PEAK*100:5.1f}%", flush=True)
    del pool, dboth
