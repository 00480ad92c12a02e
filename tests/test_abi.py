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
    for m in re.finditer(r"\b(?:int|void|size_t|uint64_t|const char\*)\s+(oct_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
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


def test_argument_validation_of_the_newer_entry_points(lib):
    P = ctypes.c_void_p
    # pooling: row range must lie inside the sequence, C % 4 == 0
    assert lib.oct_mean_pool_ws_bytes(2, 1024, 1, 5121) == 2 * 80 * 1024 * 4          # 64-row chunks, fp32 partial sums
    rc = lib.oct_mean_pool_fwd(P(16), 1, P(16), 0, 2, 10, 64, 3, 3, P(16), 1 << 20, None)
    assert rc == -1 and "row0 < row1" in _lib.last_error()
    rc = lib.oct_mean_pool_fwd(P(16), 1, P(16), 0, 2, 10, 64, 1, 10, P(16), 8, None)
    assert rc == -3 and "workspace" in _lib.last_error()
    rc = lib.oct_mean_pool_bwd(P(16), 0, P(16), 1, 2, 10, 66, 1, 10, None)
    assert rc == -1 and "C%4" in _lib.last_error()
    # ingest: W % 4, alignment, divisor
    rc = lib.oct_ingest_u8(P(16), P(16), None, None, 2, 10, 12, 64, 62, 255.0, None)
    assert rc == -1 and "W%4" in _lib.last_error()
    rc = lib.oct_ingest_u8(P(18), P(16), None, None, 2, 10, 12, 64, 64, 255.0, None)
    assert rc == -1 and "misaligned" in _lib.last_error()
    rc = lib.oct_ingest_u8(P(16), P(16), None, None, 2, 10, 12, 64, 64, 0.0, None)
    assert rc == -1 and "divisor" in _lib.last_error()
    # optimizer clock: schedule sanity, alignment
    rc = lib.oct_adamw_clock_advance(P(16), 1e-3, 1e-5, 5.0, 5.0, 0.01, 0.9, 0.95, None)
    assert rc == -1 and "schedule" in _lib.last_error()
    rc = lib.oct_adamw_clock_advance(P(8), 1e-3, 1e-5, 1.0, 5.0, 0.01, 0.9, 0.95, None)
    assert rc == -1 and "aligned" in _lib.last_error()
    rc = lib.oct_adamw_step_clocked(P(16), 4, None, 1.0, 0.9, 0.95, 1e-8, 0.05, 1.0, None, None)
    assert rc == -1 and "clock" in _lib.last_error()
    # un-shuffle: a per-sample cls row (y_row0 = 1) needs the cls row to exist
    rc = lib.oct_unshuffle_fwd(P(16), 1, P(16), P(16), P(16), None, None, P(16), 2, 16, 4, 16, 32, 1, None)
    assert rc == -1 and "y_row0" in _lib.last_error()
    rc = lib.oct_unshuffle_fwd(P(16), 1, P(16), P(16), P(16), None, P(16), P(16), 2, 16, 4, 16, 32, 2, None)
    assert rc == -1 and "y_row0" in _lib.last_error()
    # loss: the 2D model's channel-last order needs T == u and no frame index
    rc = lib.oct_mse_loss_fwd(P(16), None, P(16), 0, P(16), P(16), P(16), P(16), P(16), 2, 12, 12, 64, 64, 16, 3, 65, 1, 2, None)
    assert rc == -1 and "channel-last" in _lib.last_error()


def test_compute_entry_points_fail_loudly_without_a_device(lib):
    """No GPU (this container): a compute call returns an error code and a message — it neither crashes nor computes."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    P = ctypes.c_void_p
    rc = lib.oct_gemm(1, 0, P(1 << 20), P(2 << 20), P(3 << 20), 1, 256, 256, 256, 256, 256, 256, 0, None, None, 0, None)
    assert rc != 0 and _lib.last_error()
    rc = lib.oct_attn_fwd(1, P(1 << 20), P(2 << 20), P(3 << 20), 2, 64, 2, 32, 0.17, None)
    assert rc != 0 and _lib.last_error()
    rc = lib.oct_add_ln_fwd(P(1 << 20), 1, None, None, P(2 << 20), P(3 << 20), P(4 << 20), 1, P(5 << 20), P(6 << 20), 16, 64, 1e-6, None)
    assert rc != 0 and "oct_add_ln_fwd" in _lib.last_error()
    rc = lib.oct_ingest_u8(P(1 << 20), P(2 << 20), None, None, 2, 10, 12, 64, 64, 255.0, None)
    assert rc != 0 and "oct_ingest_u8" in _lib.last_error()


def test_no_cpu_fallback():
    """Product ops refuse CPU tensors instead of silently computing with torch."""
    import torch
    from octcubem_b200 import ops
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.mask_sort(torch.rand(2, 16), 4)
