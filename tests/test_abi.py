"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol the header declares, the ctypes
table matches the header's argument counts, and the product refuses to run without CUDA (no silent fallback)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "vaemdl.h")


def header_prototypes():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = {}
    for m in re.finditer(r"VAEMDL_API\s+([\w\s\*]+?)\s*(vaemdl_\w+)\s*\(([^;]*?)\)\s*;", text, flags=re.S):
        args = m.group(3).strip()
        n = 0 if args in ("", "void") else len([a for a in args.split(",") if a.strip()])
        protos[m.group(2)] = n
    return protos


def test_header_declares_the_expected_entry_points():
    protos = header_prototypes()
    for name in ["vaemdl_modl_fwd", "vaemdl_modl_bwd", "vaemdl_dlogistic_fwd", "vaemdl_dlogistic_bwd",
                 "vaemdl_logmeanexp_fwd", "vaemdl_logmeanexp_bwd", "vaemdl_iwae_tail", "vaemdl_modl_sample",
                 "vaemdl_dlogistic_sample", "vaemdl_modl_iwae_step_host", "vaemdl_version", "vaemdl_strerror"]:
        assert name in protos, name


def test_library_exports_every_declared_symbol(built_lib):
    for name in header_prototypes():
        assert hasattr(built_lib, name), f"{name} declared in include/vaemdl.h but not exported"


def test_ctypes_table_matches_header(built_lib):
    from vae_mdl_b200 import _abi
    protos = header_prototypes()
    assert set(protos) == set(_abi.PROTOTYPES), set(protos) ^ set(_abi.PROTOTYPES)
    for name, n_args in protos.items():
        assert len(_abi.PROTOTYPES[name][1]) == n_args, name


def test_version_and_strerror_need_no_gpu(built_lib):
    assert built_lib.vaemdl_version() == 1
    assert built_lib.vaemdl_strerror(0) == b"ok"
    assert b"aligned" in built_lib.vaemdl_strerror(-2)
    assert built_lib.vaemdl_modl_workspace_bytes(4, 32, 32) >= 4 * 1024 // 10 * 16


def test_argument_validation_without_a_gpu(built_lib):
    """Argument errors are reported before anything touches the device."""
    L = built_lib
    assert L.vaemdl_modl_fwd(None, None, 0, 0, 0, 1, 1, 4, 4, 10, None, None, None, None, 0, None) == -1
    buf = ctypes.create_string_buffer(64)
    p = ctypes.cast(buf, ctypes.c_void_p)
    assert L.vaemdl_modl_fwd(p, p, 0, 0, 0, 1, 1, 4, 4, 100, p, None, None, None, 0, None) == -3   # n_mix too large
    assert L.vaemdl_modl_fwd(p, p, 1, 1, 0, 1, 1, 4, 4, 10, p, None, None, None, 0, None) == -1    # uint8 x must be in [0,1]
    misaligned = ctypes.c_void_p(p.value + 4)
    assert L.vaemdl_modl_fwd(misaligned, p, 0, 0, 0, 1, 1, 4, 4, 10, p, None, None, None, 0, None) == -2
    assert L.vaemdl_logmeanexp_fwd(None, 1, 1, None, None) == -1


def test_product_rejects_cpu_tensors():
    """The product path must fail loudly rather than compute on the CPU."""
    import vae_mdl_b200 as V
    from vae_mdl_b200._abi import VaemdlError
    with pytest.raises(VaemdlError):
        V.MixtureDiscretizedLogistic(torch.zeros(2, 4, 4, 50))
    with pytest.raises(VaemdlError):
        V.logmeanexp(torch.zeros(5, 3), 0)
    with pytest.raises(VaemdlError):
        V.DiscretizedLogistic(torch.zeros(2, 4, 4, 3), torch.zeros(2, 4, 4, 3))
    with pytest.raises(VaemdlError):
        V.discretized_mix_logistic_loss(torch.zeros(1, 4, 4, 3), torch.zeros(1, 4, 4, 50))


def test_missing_library_raises(monkeypatch):
    from vae_mdl_b200 import _abi
    monkeypatch.setattr(_abi, "_LIB", None)
    monkeypatch.setattr(_abi, "LIB_PATH", os.path.join(ROOT, "does_not_exist.so"))
    with pytest.raises(_abi.VaemdlError, match="not found"):
        _abi.lib()


def test_product_does_not_import_the_oracle():
    """Nothing under vae_mdl_b200/ or bench's product leg may route through oracle/."""
    pkg = os.path.join(ROOT, "vae_mdl_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "import oracle" not in src and "from oracle" not in src, fn
