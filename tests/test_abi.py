"""CPU: the C-ABI library loads and exports exactly what include/octcube_b200.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

from octcubem_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "octcube_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(?:int|size_t|uint64_t|const char\*)\s+(oct_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        args = m.group(2).strip()
        n = 0 if args in ("", "void") else len([a for a in args.split(",") if a.strip()])
        out[m.group(1)] = n
    return out


@pytest.fixture(scope="module")
def lib():
    if not os.path.isfile(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    return _lib.load()


def test_header_symbols_exported(lib):
    decl = _declared()
    assert len(decl) >= 25
    for name in decl:
        assert hasattr(lib, name), f"{name} declared in the header but not exported by the .so"


def test_binding_matches_header(lib):
    decl = _declared()
    assert set(decl) == set(_lib.SIGNATURES), set(decl) ^ set(_lib.SIGNATURES)
    for name, n in decl.items():
        assert len(_lib.SIGNATURES[name][1]) == n, f"{name}: header has {n} args, binding {len(_lib.SIGNATURES[name][1])}"


def test_version_and_error_string(lib):
    assert b"octcube_b200" in lib.oct_version()
    assert isinstance(_lib.last_error(), str)


def test_argument_validation_without_gpu(lib):
    # pure host-side validation paths: must fail with OCT_ERR_INVALID before touching the device
    rc = lib.oct_mask_sort(None, 1, 8, 4, None, None, None, None)
    assert rc == -1 and "null" in _lib.last_error()
    rc = lib.oct_gemm(7, 0, ctypes.c_void_p(16), ctypes.c_void_p(16), ctypes.c_void_p(16), 0, 8, 8, 8, 8, 8, 8, 0, None, None, 0, None)
    assert rc == -1 and "compute" in _lib.last_error()
    rc = lib.oct_add_ln_fwd(ctypes.c_void_p(16), 0, None, None, ctypes.c_void_p(16), ctypes.c_void_p(16), ctypes.c_void_p(16), 0,
                            ctypes.c_void_p(16), ctypes.c_void_p(16), 4, 6, 1e-6, None)
    assert rc == -1 and "C%4" in _lib.last_error()


def test_no_cpu_fallback():
    """Product ops refuse CPU tensors instead of silently computing with torch."""
    import torch
    from octcubem_b200 import ops
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.mask_sort(torch.rand(2, 16), 4)
