"""n_mix sweep of the MoDL forward / backward kernels at a fixed ~420 MB parameter tensor, GPU only.
Prints achieved GB/s of algorithmic bytes (40*M fwd, 80*M bwd per px-sample) and writes gpurun_out/msweep.json."""
import json, os, sys
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
Ms = [int(m) for m in os.environ.get("MS", "1,2,3,4,5,6,7,8,9,10,11,12,14,15,16,18,20,24,25,30,32,40,50,64").split(",")]
out = {}
for M in Ms:
    H = W = 64; S = 16
    B = max(1, round(420e6 / (S * H * W * 40 * M)))
    p = torch.randn(S, B, H, W, 10 * M, device=dev)
    xu = torch.randint(0, 256, (B, H, W, 3), dtype=torch.uint8, device=dev)
    gimg = torch.randn(S, B, device=dev); dp = torch.empty_like(p)
    n_img = S * B; npx = n_img * H * W
    wsb = L.vaemdl_modl_workspace_bytes(n_img, H, W)
    ws = torch.empty(wsb // 8 + 1, dtype=torch.float64, device=dev); ll = torch.empty(S, B, dtype=torch.float64, device=dev)
    st = _abi.stream_ptr(dev)
    def f():
        rc = L.vaemdl_modl_fwd(p.data_ptr(), xu.data_ptr(), 1, 0, 0, n_img, B, H, W, M, None, None, ll.data_ptr(), ws.data_ptr(), wsb, st); assert rc == 0, rc
    def b():
        rc = L.vaemdl_modl_bwd(p.data_ptr(), xu.data_ptr(), 1, 0, 0, n_img, B, H, W, M, gimg.data_ptr(), None, dp.data_ptr(), st); assert rc == 0, rc
    tf_, tb = timeit(f), timeit(b)
    gf, gb = npx * 40 * M / tf_ / 1e9, npx * 80 * M / tb / 1e9
    out[M] = {"B": B, "fwd_us": tf_ * 1e6, "bwd_us": tb * 1e6, "fwd_GBs": gf, "bwd_GBs": gb}
    print(f"M={M:3d} B={B:3d}: fwd {tf_*1e6:9.1f} us {gf:7.1f} GB/s {gf/PEAK*100:5.1f}% | bwd {tb*1e6:9.1f} us {gb:7.1f} GB/s {gb/PEAK*100:5.1f}%", flush=True)
    del p, dp
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/msweep.json", "w"), indent=1)
