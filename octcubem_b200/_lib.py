"""ctypes binding of liboctcube_b200.so (the C ABI declared in include/octcube_b200.h).

There is NO fallback: if the library is missing or a call fails, a RuntimeError is raised (SURVEY §8b "Errors";
mirrors the reference's fail-fast asserts, custom_util/video_vit.py:76-78).
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_int64, c_size_t, c_void_p, POINTER

_HERE = os.path.dirname(os.path.abspath(__file__))
# OCT_LIB selects an experiment build of the same library (csrc/Makefile: BUILD= LIB= EXTRA=); never a different backend
LIB_PATH = os.environ.get("OCT_LIB") or os.path.join(_HERE, "liboctcube_b200.so")

OCT_F32, OCT_BF16, OCT_SIMT_BF16 = 0, 1, 2
GEMM_NT, GEMM_NN, GEMM_TN = 0, 1, 2
EPI_NONE, EPI_BIAS, EPI_BIAS_GELU, EPI_DGELU = 0, 1, 2, 3
LOSS_NORM_PIX, LOSS_CHANNEL_LAST, LOSS_ALL_TOKENS = 1, 2, 4

P, I, L, F, Z = c_void_p, c_int, c_int64, c_float, c_size_t

# name -> (restype, argtypes); the single source of truth for tests/test_abi.py as well
SIGNATURES = {
    "oct_version": (c_char_p, []),
    "oct_last_error": (c_char_p, []),
    "oct_launch_count": (ctypes.c_uint64, []),
    "oct_set_pdl": (None, [I]),
    "oct_device_info": (I, [POINTER(c_int), POINTER(c_int), POINTER(c_int)]),
    "oct_mask_sort": (I, [P, L, L, L, P, P, P, P]),
    "oct_patchify": (I, [P, P, I, P, P, L, L, L, L, L, L, L, L, P]),
    "oct_patch_embed_fwd": (I, [P, P, P, P, I, L, L, L, L, L, L, L, P]),
    "oct_gather_tokens_fwd": (I, [P, I, P, P, P, P, P, L, L, L, L, L, P]),
    "oct_posadd_tokens_fwd": (I, [P, I, P, P, P, P, P, L, L, L, L, L, P]),
    "oct_gather_tokens_bwd": (I, [P, P, P, I, P, P, P, L, L, L, L, L, P]),
    "oct_add_ln_fwd": (I, [P, I, P, P, P, P, P, I, P, P, L, L, F, P]),
    "oct_add_ln_bwd_ws_bytes": (Z, [L, L]),
    "oct_add_ln_bwd": (I, [P, I, P, I, P, P, P, P, P, P, I, P, P, P, Z, L, L, P]),
    "oct_add_ln_bwd_main": (I, [P, I, P, I, P, P, P, P, P, P, I, P, Z, L, L, POINTER(c_int), P]),
    "oct_add_ln_bwd_finish": (I, [P, I, L, P, P, P]),
    "oct_gemm": (I, [I, I, P, P, P, I, L, L, L, L, L, L, I, P, P, I, P]),
    "oct_gemm_wgrad_bias": (I, [I, P, P, P, P, L, L, L, L, L, L, I, P]),
    "oct_attn_fwd": (I, [I, P, P, P, L, L, L, L, F, P]),
    "oct_attn_bwd_ws_bytes": (Z, [I, L, L, L, L]),
    "oct_attn_bwd": (I, [I, P, P, P, P, P, P, Z, L, L, L, L, F, P]),
    "oct_gelu_fwd": (I, [P, P, I, L, P]),
    "oct_gelu_bwd": (I, [P, P, P, I, L, P]),
    "oct_colsum_ws_bytes": (Z, [L, L]),
    "oct_colsum": (I, [P, I, P, L, L, L, I, P, Z, P]),
    "oct_unshuffle_fwd": (I, [P, I, P, P, P, P, P, P, L, L, L, L, L, L, P]),
    "oct_unshuffle_bwd_ws_bytes": (Z, [L, L, L, L]),
    "oct_unshuffle_bwd": (I, [P, P, P, I, P, P, P, P, P, Z, L, L, L, L, L, I, L, P]),
    "oct_mse_loss_fwd": (I, [P, P, P, I, P, P, P, P, P, L, L, L, L, L, L, L, L, L, I, P]),
    "oct_mse_loss_bwd": (I, [P, P, P, I, P, P, P, P, I, L, L, L, L, L, L, L, L, L, I, P]),
    "oct_mean_pool_ws_bytes": (Z, [L, L, L, L]),
    "oct_mean_pool_fwd": (I, [P, I, P, I, L, L, L, L, L, P, Z, P]),
    "oct_mean_pool_bwd": (I, [P, I, P, I, L, L, L, L, L, P]),
    "oct_ell_spmm": (I, [P, P, P, P, L, L, L, P]),
    "oct_ingest_u8": (I, [P, P, P, P, L, L, L, L, L, F, P]),
    "oct_fg_bbox_u8": (I, [P, P, L, L, L, L, P]),
    "oct_resize_trilinear_u8": (I, [P, P, P, L, L, L, L, L, L, L, I, I, F, P]),
    "oct_cast_f32_to_bf16": (I, [P, P, L, P]),
    "oct_cast_f32_to_bf16_multi": (I, [P, L, P]),
    "oct_adamw_step": (I, [P, L, F, F, F, F, F, L, F, P, P]),
    "oct_adamw_clock_advance": (I, [P, F, F, F, F, F, F, F, P]),
    "oct_adamw_step_clocked": (I, [P, L, P, F, F, F, F, F, F, P, P]),
    "oct_grad_norm": (I, [P, L, F, F, P, P, P]),
    "oct_l2norm_fwd": (I, [P, I, P, P, L, L, F, P]),
    "oct_l2norm_bwd": (I, [P, P, P, P, I, L, L, P]),
    "oct_clip_xchg_bytes": (Z, [L, L]),
    "oct_clip_state_bytes": (Z, [L]),
    "oct_clip_loss_fwd": (I, [P, P, P, P, P, P, I, I, L, L, P]),
    "oct_clip_loss_bwd": (I, [P, P, P, P, P, P, P, P, P, I, I, L, L, P]),
    "oct_allreduce_flag_bytes": (Z, []),
    "oct_allreduce_state_bytes": (Z, []),
    "oct_allreduce_sym": (I, [P, P, L, P, L, L, I, I, I, F, I, P]),
    "oct_peer_alloc": (I, [POINTER(c_void_p), L]),
    "oct_peer_free": (I, [P]),
    "oct_peer_export": (I, [P, P]),
    "oct_peer_open": (I, [P, POINTER(c_void_p)]),
    "oct_peer_close": (I, [P]),
}

_lib = None


def load() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise RuntimeError(
                f"octcubem_b200: {LIB_PATH} not found — build it with `make -C octcubem_b200/csrc` "
                "(or `python -c 'import __graft_entry__ as g; g.build()'`). There is no CPU/PyTorch fallback."
            )
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the .so does not export what the header declares
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def last_error() -> str:
    return load().oct_last_error().decode()


def check(rc: int, what: str):
    if rc != 0:
        raise RuntimeError(f"{what} failed (code {rc}): {last_error()}")


def version() -> str:
    return load().oct_version().decode()
