"""Three forward and three backward launches of the MoDL kernels for one n_mix (ncu driver), GPU only.
    MS=5 python tools/prof_m.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vae_mdl_b200 import _abi
dev = torch.device("cuda:0")
L = _abi.lib()
for M in [int(m) for m in os.environ.get("MS", "5").split(",")]:
    H = W = 64; S = 16
    B = max(1, round(420e6 / (S * H * W * 40 * M)))
    p = torch.randn(S, B, H, W, 10 * M, device=dev)
    xu = torch.randint(0, 256, (B, H, W, 3), dtype=torch.uint8, device=dev)
    gimg = torch.randn(S, B, device=dev); dp = torch.empty_like(p)
    n_img = S * B
    wsb = L.vaemdl_modl_workspace_bytes(n_img, H, W)
    ws = torch.empty(wsb // 8 + 1, dtype=torch.float64, device=dev); ll = torch.empty(S, B, dtype=torch.float64, device=dev)
    st = _abi.stream_ptr(dev)
    for _ in range(3):
        rc = L.vaemdl_modl_fwd(p.data_ptr(), xu.data_ptr(), 1, 0, 0, n_img, B, H, W, M, None, None, ll.data_ptr(), ws.data_ptr(), wsb, st); assert rc == 0, rc
    for _ in range(3):
        rc = L.vaemdl_modl_bwd(p.data_ptr(), xu.data_ptr(), 1, 0, 0, n_img, B, H, W, M, gimg.data_ptr(), None, dp.data_ptr(), st); assert rc == 0, rc
    torch.cuda.synchronize()
    print("ok", M, float(ll.sum()), float(dp.abs().sum()))
