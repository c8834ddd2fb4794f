"""The C ABI from plain C (examples/c_abi_demo.c: no Python, no PyTorch in the process): compile with gcc, run on the GPU,
compare its loss / log-likelihoods / gradient checksums with the Python layer on the same pseudo-random inputs."""
import os
import subprocess

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "examples", "c_abi_demo.c")
EXE = os.path.join(ROOT, "examples", "c_abi_demo")


def _xorshift_stream(n):
    """The generator of examples/c_abi_demo.c (xorshift64*, top 53 bits)."""
    mask = (1 << 64) - 1
    s = 88172645463325252
    out = np.empty(n, dtype=np.float64)
    for i in range(n):
        s ^= s >> 12
        s ^= (s << 25) & mask
        s ^= s >> 27
        out[i] = ((s * 2685821657736338717) & mask) >> 11
    return out / 9007199254740992.0


def test_plain_c_program_matches_the_python_layer(built_lib):
    import vae_mdl_b200 as V
    from vae_mdl_b200 import functional as F
    if not os.path.exists(EXE) or os.path.getmtime(EXE) < os.path.getmtime(SRC):
        cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
        subprocess.run(["gcc", "-O2", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(cuda, "include"), SRC, "-o", EXE,
                        "-L", os.path.join(ROOT, "vae_mdl_b200"), "-lvaemdl_b200", "-L", os.path.join(cuda, "lib64"), "-lcudart",
                        "-lm", "-Wl,-rpath," + os.path.join(ROOT, "vae_mdl_b200")], check=True)
    S, B, H, W, M = 3, 2, 16, 16, 10
    r = subprocess.run([EXE, str(S), str(B), str(H), str(W), str(M)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    got = {ln.split()[0]: ln.split()[1:] for ln in r.stdout.strip().splitlines()}
    n_param, n_x, n_img = S * B * H * W * 10 * M, B * H * W * 3, S * B
    u = _xorshift_stream(n_param + n_x + n_img)
    params = torch.from_numpy((4.0 * u[:n_param] - 2.0).astype(np.float32)).reshape(S, B, H, W, 10 * M).cuda()
    x = torch.from_numpy((256.0 * u[n_param:n_param + n_x]).astype(np.uint8)).reshape(B, H, W, 3).cuda()
    extra = torch.from_numpy((2.0 * u[n_param + n_x:] - 1.0).astype(np.float32)).reshape(S, B).cuda()
    ll64, _, _, elbo, _, dp, launches = F.modl_iwae_step(params, x, extra)
    assert int(got["abi"][-1]) == launches          # "abi <v> S .. M .. launches <n>"
    assert abs(float(got["loss"][0]) + elbo.item()) <= 1e-6 * abs(elbo.item())
    assert abs(float(got["ll_first"][0]) - ll64.flatten()[0].item()) <= 1e-9 * abs(ll64.flatten()[0].item())
    assert abs(float(got["ll_last"][0]) - ll64.flatten()[-1].item()) <= 1e-9 * abs(ll64.flatten()[-1].item())
    assert abs(float(got["grad_abs"][0]) - dp.double().abs().sum().item()) <= 1e-6 * dp.double().abs().sum().item()
    assert abs(float(got["grad_sum"][0]) - dp.double().sum().item()) <= 1e-6 * dp.double().abs().sum().item()
