"""Five one-launch plain-DL steps at BASELINE configs[1] (ncu driver), GPU only."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vae_mdl_b200 import functional as F
dev = torch.device("cuda:0")
S, B, H, W = 5, 128, 32, 32
both = torch.randn(S, B, H, W, 6, device=dev); both[..., :3].uniform_()
x = torch.randint(0, 256, (B, H, W, 3), dtype=torch.uint8, device=dev)
for _ in range(5):
    out = F.dlogistic_iwae_step(both[..., :3], both[..., 3:], x, None, 0.0, 1.0, 256.0)
torch.cuda.synchronize()
print("launches", out[-1], float(out[3]))
