"""Plain discretized-logistic step (config 2 shape) a few times -- a short target for ncu (GPU only)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vae_mdl_b200 import _abi
dev = torch.device("cuda:0")
L = _abi.lib()
S, B, H, W = 5, 128, 32, 32
n_iter = int(sys.argv[1]) if len(sys.argv) > 1 else 6
both = torch.randn(S, B, H, W, 6, device=dev); both[..., :3].uniform_()
dboth = torch.empty_like(both)
x = torch.randint(0, 256, (B, H, W, 3), dtype=torch.uint8, device=dev)
D = H * W * 3
ll64 = torch.empty(S, B, dtype=torch.float64, device=dev); g_ll = torch.empty(S, B, device=dev)
lme = torch.empty(B, device=dev); elbo = torch.empty(1, device=dev)
wsb = L.vaemdl_dlogistic_workspace_bytes(S * B, D)
ws = torch.empty(wsb // 8 + 1, dtype=torch.float64, device=dev)
st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
p_loc, p_ls = both.data_ptr(), both.data_ptr() + 12
for _ in range(n_iter):
    rc = L.vaemdl_dlogistic_iwae_fwd(p_loc, p_ls, 3, 6, x.data_ptr(), 1, S, B, 0, B, D, 0.0, 1.0, 256.0, None, None,
                                     ll64.data_ptr(), None, lme.data_ptr(), elbo.data_ptr(), g_ll.data_ptr(), ws.data_ptr(), wsb, st)
    rc |= L.vaemdl_dlogistic_bwd(p_loc, p_ls, 3, 6, x.data_ptr(), 1, S * B, B, D, 0.0, 1.0, 256.0, g_ll.data_ptr(), None,
                                 dboth.data_ptr(), dboth.data_ptr() + 12, 6, st)
    assert rc == 0
torch.cuda.synchronize()
print("elbo", elbo.item())
