"""One-launch (cooperative) vs three-launch IWAE step, per workload, GPU only: VAEMDL_FUSED=1 / 0 through vaemdl_modl_iwae_step."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import ModlStep, WORKLOADS
dev = torch.device("cuda:0")
names = sys.argv[1:] or ["cfg1", "cfg1_m5", "cfg5_64_m10", "cfg5_64_m30", "cfg5_128_m10"]
def timeit(fn, iters=50, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3
for name in names:
    _, S, B, H, W, M = WORKLOADS[name]
    nbuf = max(2, -(-3 * 126 * 2**20 // (S * B * H * W * 40 * M)))
    st = ModlStep(S, B, H, W, M, dev, 1, B, n_buffers=nbuf)
    L = st.L
    nl = ctypes.c_int(0)
    def step_api():
        st.next_input()
        rc = L.vaemdl_modl_iwae_step(st.params.data_ptr(), st.x.data_ptr(), 1, 0, 0, S, B, st.b_total, B, H, W, M,
                                     st.extra.data_ptr(), None, st.ll64.data_ptr(), None, st.lme.data_ptr(), st.elbo.data_ptr(),
                                     st.g_ll.data_ptr(), st.dparams.data_ptr(), st.ws.data_ptr(), st.ws_bytes, st.st, ctypes.byref(nl))
        assert rc == 0, rc
    res = {}
    for mode in ("0", "1"):
        os.environ["VAEMDL_FUSED"] = mode
        t = timeit(step_api)
        res[mode] = (t, nl.value)
    gb = S * B * H * W * 120 * M / 1e9
    print(f"{name}: 3 launches {res['0'][0]:.1f} us ({gb/res['0'][0]*1e6/6549.1*100:.1f}% of peak) | "
          f"{res['1'][1]} launch {res['1'][0]:.1f} us ({gb/res['1'][0]*1e6/6549.1*100:.1f}%)", flush=True)
    del st
