"""Tuning sweep over kernel launch shapes (VAEMDL_TUNE), GPU only. Prints achieved GB/s of algorithmic bytes."""
import ctypes, json, os, sys
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
shapes_f = os.environ.get("SHAPES_F", "2:8,2:6,1:16,1:12,1:8").split(",")
shapes_b = os.environ.get("SHAPES_B", "2:8,2:6,1:16,1:14,1:12,1:8").split(",")
wls = {"cfg5a": (16, 32, 64, 64, 10), "cfg1": (5, 64, 32, 32, 10), "m30": (16, 32, 64, 64, 30), "m5": (16, 64, 64, 64, 5), "m20": (16, 32, 64, 64, 20)}
only = os.environ.get("WLS")
out = {}
for name, (S, B, H, W, M) in wls.items():
    if only and name not in only.split(","): continue
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
    for sh in shapes_f:
        os.environ["VAEMDL_TUNE"] = f"fwd={sh}"
        t = timeit(f); g = npx * 40 * M / t / 1e9
        out[f"{name} fwd {sh}"] = g
        print(f"{name:6s} fwd {sh:5s} {t*1e6:9.1f} us {g:8.1f} GB/s {g/PEAK*100:5.1f}%", flush=True)
    for sh in shapes_b:
        os.environ["VAEMDL_TUNE"] = f"bwd={sh}"
        t = timeit(b); g = npx * 80 * M / t / 1e9
        out[f"{name} bwd {sh}"] = g
        print(f"{name:6s} bwd {sh:5s} {t*1e6:9.1f} us {g:8.1f} GB/s {g/PEAK*100:5.1f}%", flush=True)
    del p, dp
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/tune.json", "w"), indent=1)
