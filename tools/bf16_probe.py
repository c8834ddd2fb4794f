"""bfloat16-parameter MoDL forward / backward vs the float32 kernels at the headline shape, GPU only."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vae_mdl_b200 import _abi
dev = torch.device("cuda:0")
L = _abi.lib()
PEAK = 6549.1
def timeit(fn, iters=30, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3
for (S, B, H, W, M) in [(16, 32, 64, 64, 10), (16, 32, 64, 64, 30), (16, 64, 64, 64, 5), (16, 20, 64, 64, 16)]:
    n_img = S * B; npx = n_img * H * W
    xu = torch.randint(0, 256, (B, H, W, 3), dtype=torch.uint8, device=dev)
    gimg = torch.randn(S, B, device=dev)
    wsb = L.vaemdl_modl_workspace_bytes(n_img, H, W)
    ws = torch.empty(wsb // 8 + 1, dtype=torch.float64, device=dev); ll = torch.empty(S, B, dtype=torch.float64, device=dev)
    st = _abi.stream_ptr(dev)
    for name, dt, fwd, bwd, bpe in (("f32 ", torch.float32, L.vaemdl_modl_fwd, L.vaemdl_modl_bwd, 4), ("bf16", torch.bfloat16, L.vaemdl_modl_fwd_bf16, L.vaemdl_modl_bwd_bf16, 2)):
        pool = [torch.randn(S, B, H, W, 10 * M, device=dev).to(dt) for _ in range(3)]
        dp = torch.empty_like(pool[0]); k = [0]
        def f():
            k[0] += 1; p = pool[k[0] % 3]
            rc = fwd(p.data_ptr(), xu.data_ptr(), 1, 0, 0, n_img, B, H, W, M, None, None, ll.data_ptr(), ws.data_ptr(), wsb, st); assert rc == 0, rc
        def b():
            k[0] += 1; p = pool[k[0] % 3]
            rc = bwd(p.data_ptr(), xu.data_ptr(), 1, 0, 0, n_img, B, H, W, M, gimg.data_ptr(), None, dp.data_ptr(), st); assert rc == 0, rc
        tf_, tb = timeit(f), timeit(b)
        gf, gb = npx * 10 * M * bpe / tf_ / 1e9, npx * 20 * M * bpe / tb / 1e9
        print(f"{S}x{B}x{H}x{W} M={M} {name}: fwd {tf_*1e6:7.1f} us {npx/tf_/1e9:6.2f} Gpx/s {gf/PEAK*100:5.1f}% | bwd {tb*1e6:7.1f} us {npx/tb/1e9:6.2f} Gpx/s {gb/PEAK*100:5.1f}%", flush=True)
        del pool, dp
