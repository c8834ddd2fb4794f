"""Host-side cost of the Python layer: vae_mdl_b200.modl_iwae_step / iwae_loss+backward vs the raw C-ABI step, GPU only."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributions as td
import vae_mdl_b200 as V
from bench import ModlStep
dev = torch.device("cuda:0")
for (S, B, H, W, M) in [(5, 64, 32, 32, 10), (5, 128, 32, 32, 5), (16, 32, 64, 64, 10)]:
    st = ModlStep(S, B, H, W, M, dev, 1, B, n_buffers=2)
    x = st.x; extra = st.extra
    def raw(): st.step()
    def py(): V.modl_iwae_step(st.pool[0], x, extra)
    z = torch.randn(S, B, 20, device=dev)
    pz = td.Normal(torch.zeros_like(z), torch.ones_like(z)); pz.axes = [-1]
    qzx = td.Normal(torch.randn(B, 20, device=dev), torch.rand(B, 20, device=dev) + 0.5); qzx.axes = [-1]
    leaf = st.pool[0].clone().requires_grad_(True)
    xf = V.normalize(x)
    def auto():
        leaf.grad = None
        loss, _ = V.iwae_loss(xf, z, pz, qzx, V.MixtureDiscretizedLogistic(leaf))
        loss.backward()
    for name, fn in (("raw C ABI step", raw), ("V.modl_iwae_step", py), ("iwae_loss + backward (autograd)", auto)):
        for _ in range(10): fn()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        n = 200
        for _ in range(n): fn()
        torch.cuda.synchronize(); t = (time.perf_counter() - t0) / n
        print(f"{S}x{B}x{H}x{W} M={M}: {name:34s} {t*1e6:8.1f} us / call (wall)", flush=True)
    del st
