"""Sweep of the run-time tiled kernel's geometry (VAEMDL_RT=LPP:rot) for a list of n_mix, GPU only."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vae_mdl_b200 import _abi
dev = torch.device("cuda:0")
L = _abi.lib()
PEAK = 6549.1
def timeit(fn, iters=10, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3
Ms = [int(m) for m in os.environ.get("MS", "11,12,16,18,25,40,50").split(",")]
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
    os.environ.pop("VAEMDL_RT", None)
    tf0, tb0 = timeit(f), timeit(b)
    print(f"M={M:3d} default: fwd {npx*40*M/tf0/1e9/PEAK*100:5.1f}% bwd {npx*80*M/tb0/1e9/PEAK*100:5.1f}%", flush=True)
    res = []
    for LPP in range(1, 17):
        MC = -(-M // LPP)
        if MC > 16 or MC < 3: continue
        for rot in (0, 1, 2, 3):
            os.environ["VAEMDL_RT"] = f"{LPP}:{rot}"
            try:
                tf_, tb = timeit(f), timeit(b)
            except AssertionError:
                continue
            res.append((npx*40*M/tf_/1e9/PEAK*100, npx*80*M/tb/1e9/PEAK*100, LPP, rot))
    res.sort(key=lambda r: -(r[0] + 2 * r[1]))
    for r in res[:6]:
        print(f"   LPP={r[2]:2d} rot={r[3]}: fwd {r[0]:5.1f}% bwd {r[1]:5.1f}%", flush=True)
    out[M] = res
    del p, dp
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/rt_sweep.json", "w"))
