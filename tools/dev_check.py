"""Developer scratch check (GPU): parity vs the oracle on small cases + quick timings. Not part of the test suite."""
import json, math, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import oracle as O
import vae_mdl_b200 as V
from vae_mdl_b200 import functional as F, _abi

dev = torch.device("cuda:0")
torch.manual_seed(0)
res = {}

def relerr(a, b):
    a = a.double().cpu(); b = b.double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-300)).item()

def modl_case(S, B, H, W, M, kind="randn", xdtype="u8"):
    g = torch.Generator().manual_seed(S * 1000 + B * 100 + M)
    xu = torch.randint(0, 256, (B, H, W, 3), dtype=torch.uint8, generator=g)
    if kind == "randn":
        p = torch.randn(S, B, H, W, 10 * M, generator=g)
    else:  # trained-like: narrow scales
        p = torch.randn(S, B, H, W, 10 * M, generator=g)
        p5 = p.view(S, B, H, W, 10, M) if False else None
        pr = p[..., M:].reshape(S, B, H, W, 3, 3 * M)
        pr[..., :M] = torch.rand(S, B, H, W, 3, M, generator=g) * 2 - 1
        pr[..., M:2 * M] = torch.randn(S, B, H, W, 3, M, generator=g) - 3.0
        p = torch.cat([p[..., :M], pr.reshape(S, B, H, W, 9 * M)], -1)
    extra = torch.randn(S, B, generator=g)
    p64 = p.double().requires_grad_(True)
    x64 = O.normalize_u8(xu, torch.float64)
    lp64 = O.modl_log_prob(p64, x64)
    ll64 = lp64.sum((-1, -2, -3))
    # make log_w comparable across S so softmax path is exercised
    shift = (-ll64.detach() + ll64.detach().mean(0, keepdim=True))
    ex = (extra.double() + shift)
    loss64, met = O.iwae_loss(lp64, ex, torch.zeros_like(ex), x64.shape)
    loss64.backward()
    pd = p.to(dev)
    xd = xu.to(dev) if xdtype == "u8" else O.normalize_u8(xu).to(dev)
    ll = F.modl_log_likelihood(pd, xd)
    lp = F.modl_log_prob(pd, xd)
    loss, lpxz, dp = V.modl_iwae_step(pd, xd, ex.float().to(dev))
    # kernel-only gradient: identical upstream weights on both sides
    gw = torch.randn(S, B, generator=g)
    p64b = p.double().requires_grad_(True)
    (O.modl_log_prob(p64b, x64).sum((-1, -2, -3)) * gw.double()).sum().backward()
    dpk = F.modl_backward(pd, xd, g_image=gw.to(dev))
    out = {
        "gradk_rel": relerr(dpk, p64b.grad),
        "ll_maxrel": ((ll.cpu().double() - ll64.detach()).abs() / ll64.detach().abs()).max().item(),
        "lp_maxabs": (lp.cpu().double() - lp64.detach()[..., 0]).abs().max().item(),
        "lpsum_vs_ll": (lp.sum((-1, -2)) - ll).abs().max().item(),
        "loss_rel": abs(loss.item() - loss64.item()) / abs(loss64.item()),
        "grad_rel": relerr(dp, p64.grad),
    }
    G = p64.grad.reshape(S, B, H, W, 10, M) if False else None
    gd = dp.cpu().double(); gr = p64.grad
    names = ["logit", "muR", "sR", "kR", "muG", "sG", "kG", "muB", "sB", "kB"]
    for j, nme in enumerate(names):
        out["g_" + nme] = relerr(gd[..., j * M:(j + 1) * M], gr[..., j * M:(j + 1) * M])
    return out

for (S, B, H, W, M, kind, xd) in [(3, 4, 8, 8, 10, "randn", "u8"), (3, 4, 8, 8, 5, "randn", "f32"), (2, 3, 8, 8, 20, "randn", "u8"),
                                  (2, 3, 8, 8, 30, "randn", "u8"), (2, 3, 5, 7, 7, "randn", "u8"), (2, 3, 5, 7, 10, "randn", "u8"),
                                  (3, 4, 8, 8, 10, "trained", "u8"), (2, 2, 32, 32, 10, "randn", "u8"), (2, 2, 32, 32, 30, "trained", "u8")]:
    key = f"modl S{S} B{B} {H}x{W} M{M} {kind} {xd}"
    try:
        res[key] = modl_case(S, B, H, W, M, kind, xd)
    except Exception as e:  # noqa
        res[key] = "ERROR " + repr(e)
    print(key, res[key], flush=True)

# plain DL
g = torch.Generator().manual_seed(5)
S, B, H, W = 3, 4, 8, 8
both = torch.randn(S, B, H, W, 6, generator=g)
xu = torch.randint(0, 256, (B, H, W, 3), dtype=torch.uint8, generator=g)
loc64 = both[..., :3].double().requires_grad_(True); ls64 = both[..., 3:].double().requires_grad_(True)
x64 = O.normalize_u8(xu, torch.float64)
lp64 = O.dlogistic_log_prob(x64, loc64, ls64, 0.0, 1.0, 256.0)
w = torch.randn(S, B, generator=g).double()
(lp64.sum((-1, -2, -3)) * w).sum().backward()
bd = both.to(dev).requires_grad_(True)
loc, ls = torch.split(bd, 3, dim=-1)
d = V.DiscretizedLogistic(loc, ls, low=0.0, high=1.0, levels=256.0)
ll = d.log_likelihood(xu.to(dev))
(ll * w.float().to(dev)).sum().backward()
lp = d.log_prob(O.normalize_u8(xu).to(dev))
res["dl"] = {"lp_maxabs": (lp.detach().cpu().double() - lp64.detach()).abs().max().item(),
             "ll_maxrel": ((ll.detach().cpu().double() - lp64.detach().sum((-1, -2, -3))).abs() / lp64.detach().sum((-1, -2, -3)).abs()).max().item(),
             "gloc": relerr(bd.grad[..., :3], loc64.grad), "gls": relerr(bd.grad[..., 3:], ls64.grad)}
print("dl", res["dl"], flush=True)

# logmeanexp
lw = torch.randn(7, 33, generator=g) * 5
lw64 = lw.double().requires_grad_(True)
o64 = O.logmeanexp(lw64, 0); o64.sum().backward()
lwd = lw.to(dev).requires_grad_(True)
o = V.logmeanexp(lwd, 0); o.sum().backward()
res["lme"] = {"fwd": (o.detach().cpu().double() - o64.detach()).abs().max().item(), "bwd": relerr(lwd.grad, lw64.grad)}
print("lme", res["lme"], flush=True)

# sampler
N, H, W, M = 6, 8, 8, 10
l = torch.randn(N, H, W, 10 * M, generator=g)
um = torch.rand(N, H, W, M, generator=g) * (1 - 2e-5) + 1e-5
ul = torch.rand(N, H, W, 3, generator=g) * (1 - 2e-5) + 1e-5
x64, idx64 = O.sample_from_discretized_mix_logistic(l, M, um, ul)
xg, xq, idx = V.sample_from_discretized_mix_logistic(l.to(dev), M, um.to(dev), ul.to(dev), return_index=True, return_quantised=True)
res["sample_openai"] = {"idx_mismatch": int((idx.cpu().long() != idx64).sum()), "x_maxabs": (xg.cpu().double() - x64).abs().max().item(),
                        "q_mismatch": int((xq.cpu() != O.quantise(x64 * 0.5 + 0.5)).sum())}
print("sample", res["sample_openai"], flush=True)
ulm = torch.rand(2, N, H, W, 3, M, generator=g) * (1 - 2e-5) + 1e-5
umm = torch.rand(2, N, H, W, M, generator=g) * (1 - 2e-5) + 1e-5
x64, idx64 = O.modl_sample_mdl(l.unsqueeze(0).expand(2, *l.shape), umm, ulm)
dist = V.MixtureDiscretizedLogistic(l.to(dev))
xg, xq, idx = dist.sample(2, u_mix=umm.to(dev), u_log=ulm.to(dev), return_index=True, return_quantised=True)
res["sample_mdl"] = {"idx_mismatch": int((idx.cpu().long() != idx64).sum()), "x_maxabs": (xg.cpu().double() - x64).abs().max().item(),
                     "q_mismatch": int((xq.cpu() != O.quantise(x64)).sum())}
print("sample_mdl", res["sample_mdl"], flush=True)

# host step
import ctypes
S, B, H, W, M = 3, 10, 8, 8, 10
p = torch.randn(S, B, H, W, 10 * M, generator=g).pin_memory()
xu = torch.randint(0, 256, (B, H, W, 3), dtype=torch.uint8, generator=g).pin_memory()
ex = torch.randn(S, B, generator=g).pin_memory()
dp = torch.empty_like(p).pin_memory(); llh = torch.empty(S, B).pin_memory(); lme = torch.empty(B).pin_memory(); elbo = torch.empty(1).pin_memory()
rc = _abi.lib().vaemdl_modl_iwae_step_host(p.data_ptr(), xu.data_ptr(), ex.data_ptr(), S, B, H, W, M, dp.data_ptr(), llh.data_ptr(), lme.data_ptr(), elbo.data_ptr(), 3)
loss_d, lpxz_d, dp_d = V.modl_iwae_step(p.to(dev), xu.to(dev), ex.to(dev))
res["host_step"] = {"rc": rc, "elbo_diff": abs(-elbo.item() - loss_d.item()), "ll_diff": (llh - lpxz_d.cpu()).abs().max().item(), "grad_diff": (dp - dp_d.cpu()).abs().max().item()}
print("host", res["host_step"], flush=True)

# timings
def timeit(fn, iters=20, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3

for name, (S, B, H, W, M) in {"cfg1": (5, 64, 32, 32, 10), "cfg1p": (5, 128, 32, 32, 5), "cfg5a": (16, 32, 64, 64, 10), "cfg5b": (16, 32, 64, 64, 30),
                              "cfg5c": (16, 32, 128, 128, 10), "m20": (16, 32, 64, 64, 20)}.items():
    p = torch.randn(S, B, H, W, 10 * M, device=dev)
    xu = torch.randint(0, 256, (B, H, W, 3), dtype=torch.uint8, device=dev)
    gimg = torch.randn(S, B, device=dev)
    dp = torch.empty_like(p)
    L = _abi.lib()
    n_img = S * B
    wsb = L.vaemdl_modl_workspace_bytes(n_img, H, W)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev); ll = torch.empty(S, B, device=dev)
    st = _abi.stream_ptr(dev)
    f = lambda: L.vaemdl_modl_fwd(p.data_ptr(), xu.data_ptr(), 1, 0, 0, n_img, B, H, W, M, None, ll.data_ptr(), None, ws.data_ptr(), wsb, st)
    b = lambda: L.vaemdl_modl_bwd(p.data_ptr(), xu.data_ptr(), 1, 0, 0, n_img, B, H, W, M, gimg.data_ptr(), None, dp.data_ptr(), st)
    tf = timeit(f); tb = timeit(b)
    npx = S * B * H * W
    res["time_" + name] = {"fwd_us": tf * 1e6, "bwd_us": tb * 1e6, "fwd_GBs": npx * 40 * M / tf / 1e9, "bwd_GBs": npx * 80 * M / tb / 1e9,
                           "fwdbwd_Gpx": npx / (tf + tb) / 1e9, "fwdbwd_GBs": npx * 120 * M / (tf + tb) / 1e9}
    print(name, res["time_" + name], flush=True)
    del p, dp

os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/dev_check.json", "w"), indent=1)
