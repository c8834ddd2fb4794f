"""Where the fwd -> finish -> bwd step spends its time (GPU only): each kernel group back to back with itself vs the step."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import ModlStep, WORKLOADS
dev = torch.device("cuda:0")
name = sys.argv[1] if len(sys.argv) > 1 else "cfg5_64_m10"
if "," in name:  # "S,B,H,W,M"
    S, B, H, W, M = (int(v) for v in name.split(","))
else:
    _, S, B, H, W, M = WORKLOADS[name]
nbuf = max(2, -(-3 * 126 * 2**20 // (S * B * H * W * 40 * M)))
st = ModlStep(S, B, H, W, M, dev, 1, B, n_buffers=nbuf)
L = st.L
def fwd_only():
    rc = L.vaemdl_modl_fwd(st.params.data_ptr(), st.x.data_ptr(), 1, 0, 0, S * B, B, H, W, M, None, None, st.ll64.data_ptr(),
                           st.ws.data_ptr(), st.ws_bytes, st.st)
    assert rc == 0
def timeit(fn, iters=50, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3
def rot(fn):
    def f():
        st.next_input(); fn()
    return f
t_f = timeit(rot(fwd_only)); t_ff = timeit(rot(st.fwd)); t_b = timeit(rot(st.bwd)); t_s = timeit(st.step)
g = torch.cuda.CUDAGraph()
st.step(); torch.cuda.synchronize()
keep = st.st
with torch.cuda.graph(g):
    st.st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)  # the capture stream
    for _ in range(len(st.pool)): st.step()
st.st = keep
t_g = timeit(g.replay, iters=20) / len(st.pool)
print(f"{name}: fwd+reduce {t_f:.1f} us | fwd+finish {t_ff:.1f} us | bwd {t_b:.1f} us | step {t_s:.1f} us | "
      f"step as CUDA graph {t_g:.1f} us | sum of parts {t_ff + t_b:.1f} us")
