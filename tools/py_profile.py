"""cProfile of the autograd route (iwae_loss + backward) at BASELINE configs[0], GPU only."""
import cProfile, os, pstats, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributions as td
import vae_mdl_b200 as V
dev = torch.device("cuda:0")
S, B, H, W, M = 5, 64, 32, 32, 10
leaf = torch.randn(S, B, H, W, 10 * M, device=dev).requires_grad_(True)
x = V.normalize(torch.randint(0, 256, (B, H, W, 3), dtype=torch.uint8, device=dev))
z = torch.randn(S, B, 20, device=dev, requires_grad=True)
pz = td.Normal(torch.zeros(S, B, 20, device=dev), torch.ones(S, B, 20, device=dev)); pz.axes = [-1]
ql = torch.randn(B, 20, device=dev, requires_grad=True); qs = (torch.rand(B, 20, device=dev) + 0.5).requires_grad_(True)
def auto():
    leaf.grad = None
    qzx = td.Normal(ql, qs); qzx.axes = [-1]
    loss, _ = V.iwae_loss(x, z, pz, qzx, V.MixtureDiscretizedLogistic(leaf))
    loss.backward()
for _ in range(20): auto()
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for _ in range(300): auto()
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
