"""bfloat16 vs float32 parameters, whole IWAE step (forward + finish + gradient), rotating inputs, at the headline shape and
its n_mix 20 / 30 siblings.  GPU only."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vae_mdl_b200 import _abi
dev = torch.device("cuda:0")
L = _abi.lib()
def timeit(fn, iters=40, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3
tag = " ".join(f"{k}={v}" for k, v in sorted(os.environ.items()) if k.startswith("VAEMDL_"))
for (S, B, H, W, M) in [(16, 32, 64, 64, 10), (16, 32, 64, 64, 20), (16, 32, 64, 64, 30)]:
    n_img = S * B
    xu = torch.randint(0, 256, (B, H, W, 3), dtype=torch.uint8, device=dev)
    extra = torch.randn(S, B, device=dev)
    wsb = L.vaemdl_modl_step_workspace_bytes(n_img, H, W)
    ws = torch.empty(wsb // 8 + 1, dtype=torch.float64, device=dev)
    ll = torch.empty(S, B, dtype=torch.float64, device=dev); g_ll = torch.empty(S, B, device=dev)
    lme = torch.empty(B, device=dev); elbo = torch.empty(1, device=dev)
    stats = torch.empty(n_img * H * W * 2, device=dev)
    st = _abi.stream_ptr(dev)
    res = {}
    for name, dt in (("f32", torch.float32), ("bf16", torch.bfloat16)):
        pool = [torch.randn(S, B, H, W, 10 * M, device=dev).to(dt) for _ in range(3)]
        dp = torch.empty_like(pool[0]); k = [0]
        fwd = L.vaemdl_modl_iwae_fwd_stats_bf16 if dt == torch.bfloat16 else L.vaemdl_modl_iwae_fwd_stats
        bwd = L.vaemdl_modl_bwd_stats_bf16 if dt == torch.bfloat16 else L.vaemdl_modl_bwd_stats
        def f():
            k[0] += 1; p = pool[k[0] % 3]
            rc = fwd(p.data_ptr(), xu.data_ptr(), 1, 0, 0, S, B, B, B, H, W, M, extra.data_ptr(), None, ll.data_ptr(), None,
                     lme.data_ptr(), elbo.data_ptr(), g_ll.data_ptr(), stats.data_ptr(), ws.data_ptr(), wsb, st); assert rc == 0, rc
        def b():
            p = pool[k[0] % 3]
            rc = bwd(p.data_ptr(), xu.data_ptr(), 1, 0, 0, n_img, B, H, W, M, g_ll.data_ptr(), None, stats.data_ptr(), dp.data_ptr(), st); assert rc == 0, rc
        def step():
            f(); b()
        res[name] = (timeit(f), timeit(lambda: (k.__setitem__(0, k[0] + 1), b())), timeit(step))
        del pool, dp
    print(f"[{tag}] {S}x{B}x{H}x{W} M={M}: " + " | ".join(f"{n}: fwd+finish {v[0]:.1f} bwd {v[1]:.1f} step {v[2]:.1f} us" for n, v in res.items())
          + f" | bf16 / f32 step time {res['bf16'][2] / res['f32'][2]:.3f}", flush=True)
