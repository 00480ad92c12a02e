"""Tensor-level wrappers over the C ABI + the autograd Functions the model is built from.

PyTorch is plumbing here (device memory, streams, autograd bookkeeping); every arithmetic op on the hot path is one
of the hand-written kernels in csrc/ reached through liboctcube_b200.so.  No op in this file has a torch fallback.
"""
from __future__ import annotations

import ctypes
import math
import os

import torch

_os_environ_get = os.environ.get

from . import _lib
from ._lib import (EPI_BIAS, EPI_BIAS_GELU, EPI_DGELU, EPI_NONE, GEMM_NN, GEMM_NT, GEMM_TN, OCT_BF16, OCT_F32,
                   OCT_SIMT_BF16)


# ----------------------------------------------------------------------------------------------------------------
# plumbing
# ----------------------------------------------------------------------------------------------------------------
def _p(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _dt(t):
    if t.dtype == torch.float32:
        return OCT_F32
    if t.dtype == torch.bfloat16:
        return OCT_BF16
    raise TypeError(f"octcubem_b200: unsupported dtype {t.dtype}")


def _chk(*ts):
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("octcubem_b200 ops need CUDA tensors (there is no CPU fallback)")
        if not t.is_contiguous():
            raise RuntimeError("octcubem_b200 ops need contiguous tensors")


def _call(name, *args):
    rc = getattr(_lib.load(), name)(*args)
    if rc != 0:
        raise RuntimeError(f"{name} failed (code {rc}): {_lib.last_error()}")


def _ws(nbytes, device):
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)


class forward_pdl:
    """Programmatic dependent launch for the kernels of a FORWARD pass (oct_set_pdl): inside the context the hot
    kernels' prologues overlap the tail of their predecessors.  Off in the backward pass, where early-resident CTAs of the dgrad
    chain would take the SMs the side-stream weight-gradient GEMMs fill.  OCT_PDL_FWD=0 disables it."""
    enabled = _os_environ_get("OCT_PDL_FWD", "0") == "1"  # measured neutral on the step (33.61 vs 33.62 ms): opt-in

    def __enter__(self):
        if forward_pdl.enabled:
            _lib.load().oct_set_pdl(1)
        return self

    def __exit__(self, *exc):
        if forward_pdl.enabled:
            _lib.load().oct_set_pdl(-1)
        return False


def compute_of(dtype):
    return OCT_F32 if dtype == torch.float32 else OCT_BF16


# ----------------------------------------------------------------------------------------------------------------
# raw ops (no autograd)
# ----------------------------------------------------------------------------------------------------------------
def mask_sort(noise: torch.Tensor, keep: int):
    """random_masking's sort part (models_mae_joint_res_flash_attn.py:356-369) -> (mask f32, ids_restore i64, ids_keep i64)."""
    _chk(noise)
    assert noise.dtype == torch.float32 and noise.dim() == 2
    B, L = noise.shape
    ids_restore = torch.empty(B, L, dtype=torch.int64, device=noise.device)
    ids_keep = torch.empty(B, keep, dtype=torch.int64, device=noise.device)
    mask = torch.empty(B, L, dtype=torch.float32, device=noise.device)
    _call("oct_mask_sort", _p(noise), B, L, keep, _p(ids_restore), _p(ids_keep), _p(mask), _stream())
    return mask, ids_restore, ids_keep


def patchify(imgs, p, u, out_dtype=torch.float32, ids_keep=None, frame_idx=None):
    _chk(imgs, ids_keep, frame_idx)
    B, C, T, H, W = imgs.shape
    assert C == 1 and imgs.dtype == torch.float32
    T_sel = T if frame_idx is None else frame_idx.numel()
    L = (T_sel // u) * (H // p) * (W // p)
    keep = 0 if ids_keep is None else ids_keep.shape[1]
    rows = L if ids_keep is None else keep
    out = torch.empty(B, rows, u * p * p, dtype=out_dtype, device=imgs.device)
    _call("oct_patchify", _p(imgs), _p(out), _dt(out), _p(ids_keep), _p(frame_idx), B, T, H, W, p, u, T_sel, keep, _stream())
    return out


def patch_embed_tc(imgs, weight2d, bias, p, u, out_dtype=torch.bfloat16):
    """Dense Conv3d patch embedding on tcgen05 (tf32), token-major output [B, L, E]."""
    _chk(imgs, weight2d, bias)
    B, C, T, H, W = imgs.shape
    assert C == 1 and imgs.dtype == torch.float32 and weight2d.dtype == torch.float32
    E = weight2d.shape[0]
    L = (T // u) * (H // p) * (W // p)
    out = torch.empty(B, L, E, dtype=out_dtype, device=imgs.device)
    _call("oct_patch_embed_fwd", _p(imgs), _p(weight2d), _p(bias), _p(out), _dt(out), B, T, H, W, p, u, E, _stream())
    return out


def ingest_u8(cube_u8, T, out=None, flip_t=None, flip_w=None, divisor=255.0):
    """uint8 cube [B, T_src, H, W] -> the step's fp32 input [B, 1, T, H, W]: /255, centre pad / crop to T frames, optional
    per-sample flips ([B] uint8 flags) — the CPU work of PatientDataset_inhouse.py:420-450 + create_3d_transforms' RandFlipd,
    bit-identical, written straight into `out` (e.g. the static input buffer of a captured step)."""
    _chk(cube_u8, out, flip_t, flip_w)
    assert cube_u8.dtype == torch.uint8 and cube_u8.dim() == 4
    B, T_src, H, W = cube_u8.shape
    if out is None:
        out = torch.empty(B, 1, T, H, W, dtype=torch.float32, device=cube_u8.device)
    assert out.dtype == torch.float32 and tuple(out.shape) == (B, 1, T, H, W)
    for f in (flip_t, flip_w):
        assert f is None or (f.dtype == torch.uint8 and f.numel() == B)
    _call("oct_ingest_u8", _p(cube_u8), _p(out), _p(flip_t), _p(flip_w), B, T_src, T, H, W, float(divisor), _stream())
    return out


def crop_resize_u8(cube_u8, T_pad, T, H, W, out=None, crop_foreground=True, flip_t=False, flip_w=False, divisor=255.0):
    """One uint8 cube [T_src, H_src, W_src] -> fp32 [T, H, W]: the loader's centre pad / crop to T_pad frames, then
    create_3d_transforms (PatientDataset_inhouse.py:56-63): CropForegroundd (bounding box of the voxels > 0; the train
    transform only), Resized(spatial_size=(T, H, W), mode="trilinear"), RandFlipd over frames / width — on the device, from the
    uint8 cube (ToTensor's /255 applied to the 8 corner voxels of every output voxel).  `out`: optional destination view."""
    _chk(cube_u8, out)
    assert cube_u8.dtype == torch.uint8 and cube_u8.dim() == 3
    T_src, Hs, Ws = cube_u8.shape
    if out is None:
        out = torch.empty(T, H, W, dtype=torch.float32, device=cube_u8.device)
    assert out.dtype == torch.float32 and out.numel() == T * H * W
    box = None
    if crop_foreground:
        box = torch.empty(6, dtype=torch.int32, device=cube_u8.device)
        _call("oct_fg_bbox_u8", _p(cube_u8), _p(box), T_src, Hs, Ws, T_pad, _stream())
    _call("oct_resize_trilinear_u8", _p(cube_u8), _p(out), _p(box), T_src, Hs, Ws, T_pad, T, H, W, int(bool(flip_t)), int(bool(flip_w)),
          float(divisor), _stream())
    return out


def gemm(layout, A, B, M, N, K, out_dtype, epilogue=EPI_NONE, bias=None, aux=None, out=None, beta=0, compute=None):
    """D[M,N] = op(A) op(B) (+ epilogue); see include/octcube_b200.h for the layouts.  A, B 2-D contiguous."""
    _chk(A, B, bias, aux, out)
    if compute is None:
        compute = compute_of(A.dtype)
    if out is None:
        out = torch.empty(M, N, dtype=out_dtype, device=A.device)
    _call("oct_gemm", compute, layout, _p(A), _p(B), _p(out), _dt(out), M, N, K, A.stride(0), B.stride(0), out.stride(0),
          epilogue, _p(bias), _p(aux), beta, _stream())
    return out


# parameter storage pointer -> (fp32 tensor that should receive that parameter's gradient, the parameter).  dp.GradReducer
# registers the views of its flat all-reduce buckets here, so that weight gradients are produced in place instead of being
# copied there.
grad_sinks = {}
_sinks_written = set()      # sinks already written during the running backward pass
_sinks_cb_queued = [False]


def _sinks_pass_done():
    _sinks_written.clear()
    _ln_seen.clear()
    _sinks_cb_queued[0] = False


def _sink(ptr, shape):
    """-> (destination or None, beta, hand_back).
    First write of a backward pass into an empty sink: beta = 0 and a FRESH alias of the sink is handed to autograd
    (AccumulateGrad adopts a gradient without copying only when nobody else holds a reference to that tensor object).
    If the sink already holds live gradient — the weight was used twice in this pass (the joint step runs a 3D and a 2D
    forward through the same blocks before one backward, engine_pretrain.py:117-149), or `param.grad` still points into the
    bucket from an earlier micro-step of a gradient-accumulation group — the kernel accumulates into it (beta = 1) and autograd
    gets None for this use: handing it a second alias of the same memory would make it add the buffer to itself."""
    ent = grad_sinks.get(ptr)
    if ent is None:
        return None, 0, True
    view, param = ent
    g = param.grad
    live = ptr in _sinks_written or (g is not None and g.data_ptr() == view.data_ptr())
    _sinks_written.add(ptr)
    _ensure_pass_callback()
    return view.view(shape), (1 if live else 0), not live


def _ensure_pass_callback():
    """Per-pass bookkeeping (_sinks_written, _ln_seen) is cleared by an autograd-engine callback at the end of the pass."""
    if not _sinks_cb_queued[0]:
        _sinks_cb_queued[0] = True
        torch.autograd.Variable._execution_engine.queue_callback(_sinks_pass_done)


# ----------------------------------------------------------------------------------------------------------------
# weight gradients on a side stream: in the backward pass only the dgrad chain is sequential; the wgrad GEMMs are leaves.
# Launched on a second stream (a parallel branch of the captured CUDA graph) their CTAs fill the SMs that the chain's
# kernels leave idle (the encoder's N = 1024 GEMMs occupy 104 of 148 SMs, LN backward is HBM-bound, ...).
# ----------------------------------------------------------------------------------------------------------------
import os as _os

overlap_wgrad = _os.environ.get("OCT_WGRAD_STREAM", "1") != "0"
_wgrad_streams = {}
_wgrad_pending = {}


def wgrad_stream(device):
    st = _wgrad_streams.get(device.index)
    if st is None:
        st = _wgrad_streams[device.index] = torch.cuda.Stream(device=device)
    return st


def _join_wgrad(device):
    """End of the backward pass (autograd engine callback): everything after it sees the weight gradients."""
    def cb():
        if _wgrad_pending.pop(device.index, False):
            torch.cuda.current_stream(device).wait_stream(wgrad_stream(device))
    return cb


def join_wgrad(device=None):
    """Make the current stream wait for the weight gradients launched so far (a no-op when nothing is pending).  The
    autograd-engine callback does this at the end of every backward pass; optim.FusedAdamW calls it defensively."""
    for idx in list(_wgrad_pending):
        if (device is None or device.index == idx) and _wgrad_pending.pop(idx, False):
            torch.cuda.current_stream(torch.device("cuda", idx)).wait_stream(_wgrad_streams[idx])


def wgrad_bias_async(dy2, x2, dw=None, db=None, beta=0):
    """wgrad_bias on the side stream (inside a backward pass, bf16 path); falls back to the current stream otherwise."""
    if not overlap_wgrad or dy2.dtype != torch.bfloat16:
        return wgrad_bias(dy2, x2, dw, db, beta)
    dev = dy2.device
    side, cur = wgrad_stream(dev), torch.cuda.current_stream(dev)
    # one join callback per call (the first to run does the work): a backward pass that died half-way leaves the flag set,
    # and the next pass must still be joined
    _wgrad_pending[dev.index] = True
    torch.autograd.Variable._execution_engine.queue_callback(_join_wgrad(dev))
    side.wait_stream(cur)  # dy2 / x2 are complete on the current stream
    with torch.cuda.stream(side):
        out = wgrad_bias(dy2, x2, dw, db, beta)
    # the operands were allocated on the current stream: the caching allocator must not recycle them under the side stream
    dy2.record_stream(side)
    x2.record_stream(side)
    return out


def wgrad_bias(dy2, x2, dw=None, db=None, beta=0):
    """(dW [n_out, k_in], db [n_out]) (= or, beta = 1, +=) (dy2^T x2, column sums of dy2), both fp32.  bf16 operands: ONE
    tcgen05 kernel (oct_gemm_wgrad_bias: the bias gradient rides on the wgrad GEMM as an extra MMA against a tile of ones);
    fp32 parity mode: CUDA-core GEMM + the deterministic two-stage column sum.  dw / db: optional destinations (grad_sinks)."""
    _chk(dy2, x2, dw, db)
    assert beta == 0 or (dw is not None and db is not None)
    tokens, n_out = dy2.shape
    k_in = x2.shape[1]
    if dy2.dtype != torch.bfloat16:
        return gemm(GEMM_TN, dy2, x2, n_out, k_in, tokens, torch.float32, out=dw, beta=beta), colsum(dy2, out=db, beta=beta)
    if dw is None:
        dw = torch.empty(n_out, k_in, dtype=torch.float32, device=dy2.device)
    if db is None:
        db = torch.empty(n_out, dtype=torch.float32, device=dy2.device)
    _call("oct_gemm_wgrad_bias", OCT_BF16, _p(dy2), _p(x2), _p(dw), _p(db), n_out, k_in, tokens, dy2.stride(0), x2.stride(0),
          dw.stride(0), beta, _stream())
    return dw, db


def colsum(x2d, out=None, beta=0):
    _chk(x2d)
    M, N = x2d.shape
    if out is None:
        out = torch.empty(N, dtype=torch.float32, device=x2d.device)
    nb = _lib.load().oct_colsum_ws_bytes(M, N)
    ws = _ws(nb, x2d.device)
    _call("oct_colsum", _p(x2d), _dt(x2d), _p(out), M, N, x2d.stride(0), beta, _p(ws), ws.numel(), _stream())
    return out


def add_ln_fwd(h, res_in, gamma, beta, eps, y_dtype, want_res_out):
    _chk(h, res_in, gamma, beta)
    C = h.shape[-1]
    M = h.numel() // C
    y = torch.empty(h.shape, dtype=y_dtype, device=h.device)
    res_out = torch.empty(h.shape, dtype=torch.float32, device=h.device) if want_res_out else None
    mean = torch.empty(M, dtype=torch.float32, device=h.device)
    rstd = torch.empty(M, dtype=torch.float32, device=h.device)
    _call("oct_add_ln_fwd", _p(h), _dt(h), _p(res_in), _p(res_out), _p(gamma), _p(beta), _p(y), _dt(y), _p(mean), _p(rstd),
          M, C, float(eps), _stream())
    return y, res_out, mean, rstd


def add_ln_bwd(dy, x, mean, rstd, gamma, dres_in, want_f32, want_lp, beta_param=None):
    """-> (dx_f32, dx_lp, dgamma, dbeta).  `beta_param` (the LayerNorm bias Parameter, optional) enables the deferred reduction:
    dx continues the dgrad chain while the fixed-order reduction of the dgamma / dbeta partials — a leaf, like the weight
    gradients — runs on their side stream, straight into the reducer's bucket views when those are registered (66 launches per
    step off the critical path).  dgamma / dbeta come back as None when they were accumulated into live gradients in place."""
    _chk(dy, x, mean, rstd, gamma, dres_in)
    C = x.shape[-1]
    M = x.numel() // C
    dx_f32 = torch.empty(x.shape, dtype=torch.float32, device=x.device) if want_f32 else None
    dx_lp = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device) if want_lp else None
    nb = _lib.load().oct_add_ln_bwd_ws_bytes(M, C)
    ws = _ws(nb, x.device)
    dev = x.device
    defer = overlap_wgrad and beta_param is not None and dy.dtype == torch.bfloat16
    if defer:
        # only when nobody on the CURRENT stream reads the results before the end-of-pass join: the first gradient of each
        # parameter in this pass, with no live .grad (autograd then adopts the tensors without touching them)
        keys = (gamma.data_ptr(), beta_param.data_ptr())
        g_dst, g_beta, g_give = _sink(keys[0], (C,))
        b_dst, b_beta, b_give = _sink(keys[1], (C,))
        fresh = all(k not in _ln_seen for k in keys) and g_beta == 0 and b_beta == 0
        if g_dst is None:   # no reducer: plain autograd gradients
            fresh = fresh and _param_grad_is_none(gamma) and _param_grad_is_none(beta_param)
        _ln_seen.update(keys)
        _ensure_pass_callback()
        if not fresh:
            defer = False
            if _wgrad_pending.get(dev.index, False):  # order behind earlier deferred writes into the same buffers
                torch.cuda.current_stream(dev).wait_stream(wgrad_stream(dev))
    if not defer:
        dgamma = torch.empty(C, dtype=torch.float32, device=dev)
        dbeta = torch.empty(C, dtype=torch.float32, device=dev)
        _call("oct_add_ln_bwd", _p(dy), _dt(dy), _p(x), _dt(x), _p(mean), _p(rstd), _p(gamma), _p(dres_in), _p(dx_f32), _p(dx_lp),
              OCT_BF16, _p(dgamma), _p(dbeta), _p(ws), ws.numel(), M, C, _stream())
        return dx_f32, dx_lp, dgamma, dbeta
    dgamma = g_dst if g_dst is not None else torch.empty(C, dtype=torch.float32, device=dev)
    dbeta = b_dst if b_dst is not None else torch.empty(C, dtype=torch.float32, device=dev)
    nblocks = ctypes.c_int(0)
    _call("oct_add_ln_bwd_main", _p(dy), _dt(dy), _p(x), _dt(x), _p(mean), _p(rstd), _p(gamma), _p(dres_in), _p(dx_f32), _p(dx_lp),
          OCT_BF16, _p(ws), ws.numel(), M, C, ctypes.byref(nblocks), _stream())
    side, cur = wgrad_stream(dev), torch.cuda.current_stream(dev)
    _wgrad_pending[dev.index] = True
    torch.autograd.Variable._execution_engine.queue_callback(_join_wgrad(dev))
    side.wait_stream(cur)
    with torch.cuda.stream(side):
        _call("oct_add_ln_bwd_finish", _p(ws), nblocks.value, C, _p(dgamma), _p(dbeta), _stream())
    for t in (ws, dgamma, dbeta):
        t.record_stream(side)
    return dx_f32, dx_lp, dgamma, dbeta


_ln_seen = set()      # LayerNorm parameters that already received a gradient in the running backward pass
_param_by_ptr = {}    # storage pointer -> LayerNorm Parameter (registered by AddLNFn.forward; weak by construction: ptr reuse is harmless,
                      # a stale entry can only make a step take the synchronous path)


def _param_grad_is_none(t):
    p = _param_by_ptr.get(t.data_ptr())
    return p is not None and p.grad is None


def attn_fwd(qkv, H, d, compute):
    _chk(qkv)
    B, S = qkv.shape[0], qkv.shape[1]
    out = torch.empty(B, S, H * d, dtype=qkv.dtype, device=qkv.device)
    lse = torch.empty(B, H, S, dtype=torch.float32, device=qkv.device)
    _call("oct_attn_fwd", compute, _p(qkv), _p(out), _p(lse), B, S, H, d, 1.0 / math.sqrt(d), _stream())
    return out, lse


def attn_bwd(qkv, out, dout, lse, H, d, compute):
    _chk(qkv, out, dout, lse)
    B, S = qkv.shape[0], qkv.shape[1]
    dqkv = torch.empty_like(qkv)
    nb = _lib.load().oct_attn_bwd_ws_bytes(compute, B, S, H, d)
    ws = _ws(nb, qkv.device)
    _call("oct_attn_bwd", compute, _p(qkv), _p(out), _p(dout), _p(lse), _p(dqkv), _p(ws), ws.numel(), B, S, H, d,
          1.0 / math.sqrt(d), _stream())
    return dqkv


def gelu_fwd(x):
    _chk(x)
    y = torch.empty_like(x)
    _call("oct_gelu_fwd", _p(x), _p(y), _dt(x), x.numel(), _stream())
    return y


def gelu_bwd(dy, x):
    _chk(dy, x)
    dx = torch.empty_like(x)
    _call("oct_gelu_bwd", _p(dy), _p(x), _p(dx), _dt(x), x.numel(), _stream())
    return dx


def cast_bf16(src_f32, dst_bf16=None):
    _chk(src_f32, dst_bf16)
    if dst_bf16 is None:
        dst_bf16 = torch.empty(src_f32.shape, dtype=torch.bfloat16, device=src_f32.device)
    _call("oct_cast_f32_to_bf16", _p(src_f32), _p(dst_bf16), src_f32.numel(), _stream())
    return dst_bf16


CAST_CHUNK = 16384


def cast_table(pairs):
    """Device chunk table for cast_bf16_multi: `pairs` = [(fp32 source, bf16 destination)], same numel, contiguous."""
    rows = []
    for src, dst in pairs:
        _chk(src, dst)
        assert src.dtype == torch.float32 and dst.dtype == torch.bfloat16 and src.numel() == dst.numel()
        assert src.is_contiguous() and dst.is_contiguous() and src.data_ptr() % 16 == 0 and dst.data_ptr() % 8 == 0
        n = src.numel()
        for off in range(0, n, CAST_CHUNK):
            rows.append((src.data_ptr() + 4 * off, dst.data_ptr() + 2 * off, min(CAST_CHUNK, n - off)))
    dev = pairs[0][0].device
    return torch.tensor(rows, dtype=torch.int64).to(dev), len(rows)


def cast_bf16_multi(table, n_chunks):
    """fp32 -> bf16 of every (source, destination) pair of a cast_table in one launch."""
    _call("oct_cast_f32_to_bf16_multi", _p(table), n_chunks, _stream())


def _wgrad_into_sinks(dy2, x2, w_sink, b_sink):
    """Weight + bias gradient of one nn.Linear, produced in the reducer's buckets when they are registered (see _sink)."""
    dw, beta_w, give_w = _sink(*w_sink)
    db, beta_b, give_b = _sink(*b_sink)
    if beta_w != beta_b or (dw is None) != (db is None):  # the pair always travels together; be safe if it ever does not
        dw2, db2 = wgrad_bias_async(dy2, x2)
        return dw2, db2
    dw, db = wgrad_bias_async(dy2, x2, dw, db, beta_w)
    return (dw if give_w else None), (db if give_b else None)


# ----------------------------------------------------------------------------------------------------------------
# autograd Functions
# ----------------------------------------------------------------------------------------------------------------
class LinearFn(torch.autograd.Function):
    """y = x W^T + b (nn.Linear; flash_attn/modules/mha.py:635,703, models:511,595).  `w_lp` is the bf16 shadow of `weight`
    used by the tensor-core path (None in fp32 mode).  Grad of weight/bias is fp32."""

    @staticmethod
    def forward(ctx, x, weight, bias, w_lp):
        w = weight if w_lp is None else w_lp
        assert x.dtype == w.dtype, (x.dtype, w.dtype)
        N, K = w.shape
        x2 = x.reshape(-1, K)
        M = x2.shape[0]
        y = gemm(GEMM_NT, x2, w, M, N, K, x.dtype, EPI_BIAS, bias=bias)
        ctx.save_for_backward(x2, w)
        ctx.xshape = x.shape
        ctx.sinks = (weight.data_ptr(), tuple(weight.shape), bias.data_ptr(), tuple(bias.shape))
        return y.view(*x.shape[:-1], N)

    @staticmethod
    def backward(ctx, dy):
        x2, w = ctx.saved_tensors
        N, K = w.shape
        dy2 = dy.reshape(-1, N)
        if not dy2.is_contiguous():
            dy2 = dy2.contiguous()
        M = x2.shape[0]
        dx = dw = db = None
        if ctx.needs_input_grad[1] and ctx.needs_input_grad[2]:  # first: it forks onto the wgrad stream and overlaps the dgrad
            wp, ws_, bp, bs = ctx.sinks
            dw, db = _wgrad_into_sinks(dy2, x2, (wp, ws_), (bp, bs))
        elif ctx.needs_input_grad[1]:
            dw = gemm(GEMM_TN, dy2, x2, N, K, M, torch.float32)
        elif ctx.needs_input_grad[2]:
            db = colsum(dy2)
        if ctx.needs_input_grad[0]:
            dx = gemm(GEMM_NN, dy2, w, M, K, N, x2.dtype).view(ctx.xshape)
        return dx, dw, db, None


class MlpFn(torch.autograd.Function):
    """fc1 -> GELU(erf) -> fc2 (flash_attn/modules/mlp.py:47-51).  GELU is fused into the fc1 epilogue and its derivative
    into the fc2-dgrad epilogue (bf16 and fp32 paths alike)."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, w1_lp, w2_lp):
        wa = w1 if w1_lp is None else w1_lp
        wb = w2 if w2_lp is None else w2_lp
        hid, dim = wa.shape
        x2 = x.reshape(-1, dim)
        M = x2.shape[0]
        pre = torch.empty(M, hid, dtype=x.dtype, device=x.device)
        act = gemm(GEMM_NT, x2, wa, M, hid, dim, x.dtype, EPI_BIAS_GELU, bias=b1, aux=pre)
        y = gemm(GEMM_NT, act, wb, M, wb.shape[0], hid, x.dtype, EPI_BIAS, bias=b2)
        ctx.save_for_backward(x2, wa, wb, pre, act)
        ctx.xshape = x.shape
        ctx.sinks = tuple((t.data_ptr(), tuple(t.shape)) for t in (w1, b1, w2, b2))
        return y.view(*x.shape[:-1], wb.shape[0])

    @staticmethod
    def backward(ctx, dy):
        x2, wa, wb, pre, act = ctx.saved_tensors
        hid, dim = wa.shape
        out_dim = wb.shape[0]
        dy2 = dy.reshape(-1, out_dim)
        if not dy2.is_contiguous():
            dy2 = dy2.contiguous()
        M = x2.shape[0]
        (w1s, b1s, w2s, b2s) = ctx.sinks
        dw2, db2 = _wgrad_into_sinks(dy2, act, w2s, b2s)  # forks onto the wgrad stream: overlaps the dgrad chain
        dpre = gemm(GEMM_NN, dy2, wb, M, hid, out_dim, x2.dtype, EPI_DGELU, aux=pre)
        dw1, db1 = _wgrad_into_sinks(dpre, x2, w1s, b1s)
        dx = gemm(GEMM_NN, dpre, wa, M, dim, hid, x2.dtype).view(ctx.xshape) if ctx.needs_input_grad[0] else None
        return dx, dw1, db1, dw2, db2, None, None


class AttnFn(torch.autograd.Function):
    """flash_attn_qkvpacked_func(qkv, 0.0, softmax_scale=d^-0.5, causal=False) (flash_attn/modules/mha.py:122-130)."""

    @staticmethod
    def forward(ctx, qkv, H, compute):
        B, S, three_dim = qkv.shape
        d = three_dim // (3 * H)
        out, lse = attn_fwd(qkv, H, d, compute)
        ctx.save_for_backward(qkv, out, lse)
        ctx.H, ctx.d, ctx.compute = H, d, compute
        return out

    @staticmethod
    def backward(ctx, dout):
        qkv, out, lse = ctx.saved_tensors
        if not dout.is_contiguous():
            dout = dout.contiguous()
        return attn_bwd(qkv, out, dout, lse, ctx.H, ctx.d, ctx.compute), None, None


class AddLNFn(torch.autograd.Function):
    """residual = h (+ residual_in);  y = LayerNorm(residual)  (flash_attn/modules/block.py:126-130,163-167).
    Returns (y, residual fp32).  With residual_in=None and fp32 h the residual IS h (block 0) and is not copied.
    keep_residual=False is the final-norm case (models:489,592): only y is produced (quirk Q1)."""

    @staticmethod
    def forward(ctx, h, res_in, gamma, beta, eps, y_dtype, keep_residual):
        if res_in is not None and not keep_residual:
            raise RuntimeError("AddLNFn: res_in given but residual not kept")
        y, res_out, mean, rstd = add_ln_fwd(h, res_in, gamma, beta, eps, y_dtype, keep_residual)
        x = res_out if keep_residual else h  # the tensor that was normalised
        if isinstance(gamma, torch.nn.Parameter) and isinstance(beta, torch.nn.Parameter):
            _param_by_ptr[gamma.data_ptr()] = gamma
            _param_by_ptr[beta.data_ptr()] = beta
            ctx.beta_param = beta
        else:
            ctx.beta_param = None
        ctx.save_for_backward(x, mean, rstd, gamma)
        ctx.h_dtype = h.dtype
        ctx.has_res_in = res_in is not None
        return y, res_out

    @staticmethod
    def backward(ctx, dy, dres):
        x, mean, rstd, gamma = ctx.saved_tensors
        if not dy.is_contiguous():
            dy = dy.contiguous()
        if dres is not None and not dres.is_contiguous():
            dres = dres.contiguous()
        h_lp = ctx.h_dtype == torch.bfloat16
        want_f32 = ctx.has_res_in or not h_lp
        dx_f32, dx_lp, dgamma, dbeta = add_ln_bwd(dy, x, mean, rstd, gamma, dres, want_f32, h_lp, ctx.beta_param)
        dh = dx_lp if h_lp else dx_f32
        dres_in = dx_f32 if ctx.has_res_in else None
        return dh, dres_in, dgamma, dbeta, None, None, None


class GatherTokensFn(torch.autograd.Function):
    """x_masked + cls + pos (models:406-478): out[b,0]=cls_row; out[b,1+i] = x[b,ids_keep[b,i]] + pos_sp[s] + pos_tmp[t].
    x is the dense patch-embed output; its gradient is returned as a dense tensor only if required (standalone use);
    the fused encoder path uses EmbedTokensFn below, which never materialises the dense gradient."""

    @staticmethod
    def forward(ctx, x, ids_keep, pos_sp, pos_tmp, cls_row):
        _chk(x, ids_keep, pos_sp, pos_tmp, cls_row)
        B, L, C = x.shape
        keep = ids_keep.shape[1]
        G = pos_sp.shape[0] if pos_sp is not None else L
        out = torch.empty(B, keep + (1 if cls_row is not None else 0), C, dtype=torch.float32, device=x.device)
        _call("oct_gather_tokens_fwd", _p(x), _dt(x), _p(ids_keep), _p(pos_sp), _p(pos_tmp), _p(cls_row), _p(out), B, L, keep,
              G, C, _stream())
        ctx.save_for_backward(ids_keep)
        ctx.dims = (B, L, keep, G, C, x.dtype, pos_sp is not None, pos_tmp is not None, cls_row is not None)
        return out

    @staticmethod
    def backward(ctx, dout):
        (ids_keep,) = ctx.saved_tensors
        B, L, keep, G, C, xdt, has_sp, has_tmp, has_cls = ctx.dims
        if not dout.is_contiguous():
            dout = dout.contiguous()
        dev = dout.device
        dxk = torch.empty(B * keep, C, dtype=xdt, device=dev)
        d_sp = torch.empty(G, C, dtype=torch.float32, device=dev) if has_sp else None
        d_tmp = torch.empty(L // G, C, dtype=torch.float32, device=dev) if has_tmp else None
        d_cls = torch.empty(C, dtype=torch.float32, device=dev) if has_cls else None
        _call("oct_gather_tokens_bwd", _p(dout), _p(ids_keep), _p(dxk), _dt(dxk), _p(d_sp), _p(d_tmp), _p(d_cls), B, L, keep, G,
              C, _stream())
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.zeros(B, L, C, dtype=xdt, device=dev)
            dx.scatter_(1, ids_keep.unsqueeze(-1).expand(-1, -1, C), dxk.view(B, keep, C))  # standalone use only
        return dx, None, d_sp, d_tmp, d_cls


class EmbedTokensFn(torch.autograd.Function):
    """Fused encoder front end: PatchEmbed (vv:74-83) -> random_masking gather (models:362-363) -> cls + pos add
    (models:409-478), GATHER FIRST: `ids_keep` depends only on the noise, so only the kept tokens (10 % at mask 0.9) are
    patchified and embedded — one [B*keep, u*p*p] x [E, u*p*p]^T GEMM with the bias in its epilogue (bf16 operands, which is
    also what the reference's Conv3d computes under autocast; fp32 in the parity mode) instead of the dense embedding of all L
    tokens followed by a gather.  The kept patches are saved: the backward's dW = dX_keep^T · patches_keep needs them again.
    `w_lp`: bf16 shadow of the Conv3d weight (None in fp32 mode)."""

    @staticmethod
    def forward(ctx, imgs, weight, bias, ids_keep, pos_sp, pos_tmp, cls_row, p, u, act_dtype, w_lp=None):
        _chk(imgs, weight, bias, ids_keep, pos_sp, pos_tmp, cls_row, w_lp)
        B, _, T, H, W = imgs.shape
        E = weight.shape[0]
        w2d = (weight if w_lp is None else w_lp).view(E, -1)
        L = (T // u) * (H // p) * (W // p)
        keep = ids_keep.shape[1]
        G = pos_sp.shape[0]
        pk = patchify(imgs, p, u, act_dtype, ids_keep=ids_keep).view(B * keep, -1)
        x = gemm(GEMM_NT, pk, w2d, B * keep, E, w2d.shape[1], act_dtype, EPI_BIAS, bias=bias)
        out = torch.empty(B, keep + (1 if cls_row is not None else 0), E, dtype=torch.float32, device=imgs.device)
        _call("oct_posadd_tokens_fwd", _p(x), _dt(x), _p(ids_keep), _p(pos_sp), _p(pos_tmp), _p(cls_row), _p(out), B, L, keep,
              G, E, _stream())
        ctx.save_for_backward(pk, ids_keep)
        ctx.dims = (B, L, keep, G, E, act_dtype, pos_tmp is not None, cls_row is not None, tuple(weight.shape))
        return out

    @staticmethod
    def backward(ctx, dout):
        pk, ids_keep = ctx.saved_tensors
        B, L, keep, G, E, act_dtype, has_tmp, has_cls, wshape = ctx.dims
        if not dout.is_contiguous():
            dout = dout.contiguous()
        dev = dout.device
        dxk = torch.empty(B * keep, E, dtype=act_dtype, device=dev)
        d_sp = torch.empty(G, E, dtype=torch.float32, device=dev)
        d_tmp = torch.empty(L // G, E, dtype=torch.float32, device=dev) if has_tmp else None
        d_cls = torch.empty(E, dtype=torch.float32, device=dev) if has_cls else None
        _call("oct_gather_tokens_bwd", _p(dout), _p(ids_keep), _p(dxk), _dt(dxk), _p(d_sp), _p(d_tmp), _p(d_cls), B, L, keep, G,
              E, _stream())
        dw, db = wgrad_bias(dxk, pk)
        return None, dw.view(wshape), db, None, d_sp, d_tmp, d_cls, None, None, None, None


class UnshuffleFn(torch.autograd.Function):
    """Decoder input assembly (models:515-573): mask tokens + un-shuffle by ids_restore + decoder cls + pos add.
    y_row0 = 1 is the 2D model's variant (OCTCube/models_mae_flash_attn.py:299-312): y [B, 1 + keep, D] carries the sample's
    own cls token in row 0 and out[b, 0] = y[b, 0] + cls_row."""

    @staticmethod
    def forward(ctx, y, ids_restore, mask_token, pos_sp, pos_tmp, cls_row, y_row0=0):
        _chk(y, ids_restore, mask_token, pos_sp, pos_tmp, cls_row)
        B, D = y.shape[0], y.shape[2]
        keep = y.shape[1] - y_row0
        L = ids_restore.shape[1]
        G = pos_sp.shape[0]
        has_cls = cls_row is not None
        out = torch.empty(B, L + (1 if has_cls else 0), D, dtype=torch.float32, device=y.device)
        _call("oct_unshuffle_fwd", _p(y), _dt(y), _p(ids_restore), _p(mask_token), _p(pos_sp), _p(pos_tmp), _p(cls_row), _p(out),
              B, L, keep, G, D, y_row0, _stream())
        ctx.save_for_backward(ids_restore)
        ctx.dims = (B, L, keep, G, D, y.dtype, pos_tmp is not None, has_cls, y_row0)
        return out

    @staticmethod
    def backward(ctx, dout):
        (ids_restore,) = ctx.saved_tensors
        B, L, keep, G, D, ydt, has_tmp, has_cls, y_row0 = ctx.dims
        if not dout.is_contiguous():
            dout = dout.contiguous()
        dev = dout.device
        dy = torch.empty(B, y_row0 + keep, D, dtype=ydt, device=dev)
        d_mt = torch.empty(D, dtype=torch.float32, device=dev)
        d_sp = torch.empty(G, D, dtype=torch.float32, device=dev)
        d_tmp = torch.empty(L // G, D, dtype=torch.float32, device=dev) if has_tmp else None
        d_cls = torch.empty(D, dtype=torch.float32, device=dev) if has_cls else None
        nb = _lib.load().oct_unshuffle_bwd_ws_bytes(B, L, G, D)
        ws = _ws(nb, dev)
        _call("oct_unshuffle_bwd", _p(dout), _p(ids_restore), _p(dy), _dt(dy), _p(d_mt), _p(d_sp), _p(d_tmp), _p(d_cls), _p(ws),
              ws.numel(), B, L, keep, G, D, 1 if has_cls else 0, y_row0, _stream())
        return dy, None, d_mt, d_sp, d_tmp, d_cls, None


class MaskedMSELossFn(torch.autograd.Function):
    """forward_loss (models:613-667).  pred_full [B, row0 + L, P] (row0 = 1: the cls row is skipped in place).
    Returns (loss, frame_losses, loss_tok); frame_losses / loss_tok carry no gradient (the engine only logs them,
    engine_pretrain.py:133-146).  `extra_flags`: LOSS_CHANNEL_LAST / LOSS_ALL_TOKENS for the 2D model
    (OCTCube/models_mae_flash_attn.py:331-350; imgs is then the [B,C,H,W] image viewed as [B,1,C,H,W], u = C)."""

    @staticmethod
    def forward(ctx, imgs, pred_full, mask, p, u, row0, norm_pix, frame_idx, extra_flags=0):
        _chk(imgs, pred_full, mask, frame_idx)
        B, _, T, H, W = imgs.shape
        T_sel = T if frame_idx is None else frame_idx.numel()
        Tp = T_sel // u
        L = mask.shape[1]
        dev = imgs.device
        loss_tok = torch.empty(B, L, dtype=torch.float32, device=dev)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        msum = torch.empty((), dtype=torch.float32, device=dev)
        frame = torch.empty(B, Tp, dtype=torch.float32, device=dev)
        flags = (_lib.LOSS_NORM_PIX if norm_pix else 0) | int(extra_flags)
        _call("oct_mse_loss_fwd", _p(imgs), _p(frame_idx), _p(pred_full), _dt(pred_full), _p(mask), _p(loss_tok), _p(loss),
              _p(msum), _p(frame), B, T, T_sel, H, W, p, u, pred_full.shape[1], row0, flags, _stream())
        ctx.save_for_backward(imgs, pred_full, mask, msum, frame_idx if frame_idx is not None else torch.empty(0, device=dev))
        ctx.cfg = (p, u, row0, flags, frame_idx is not None)
        ctx.mark_non_differentiable(frame, loss_tok)
        return loss, frame, loss_tok

    @staticmethod
    def backward(ctx, dloss, _dframe, _dtok):
        imgs, pred_full, mask, msum, frame_idx = ctx.saved_tensors
        p, u, row0, flags, has_idx = ctx.cfg
        if not has_idx:
            frame_idx = None
        B, _, T, H, W = imgs.shape
        T_sel = T if frame_idx is None else frame_idx.numel()
        dloss = dloss.to(torch.float32).contiguous()
        dpred = torch.empty_like(pred_full)
        _call("oct_mse_loss_bwd", _p(imgs), _p(frame_idx), _p(pred_full), _dt(pred_full), _p(mask), _p(msum), _p(dloss), _p(dpred),
              _dt(dpred), B, T, T_sel, H, W, p, u, pred_full.shape[1], row0, flags, _stream())
        return None, dpred, None, None, None, None, None, None, None


class MeanPoolFn(torch.autograd.Function):
    """x[:, row0:row1, :].mean(dim=1) -> [B, C] in `out_dtype` (OCTCube/models_vit_st_flash_attn.py:247-251: global pool
    without the cls token = rows [1, S); the cls read-out `x[:, 0]` = rows [0, 1))."""

    @staticmethod
    def forward(ctx, x, row0, row1, out_dtype):
        _chk(x)
        B, S, C = x.shape
        out = torch.empty(B, C, dtype=out_dtype, device=x.device)
        ws = _ws(_lib.load().oct_mean_pool_ws_bytes(B, C, row0, row1), x.device)
        _call("oct_mean_pool_fwd", _p(x), _dt(x), _p(out), _dt(out), B, S, C, row0, row1, _p(ws), ws.numel(), _stream())
        ctx.dims = (B, S, C, row0, row1, x.dtype)
        return out

    @staticmethod
    def backward(ctx, dout):
        B, S, C, row0, row1, xdt = ctx.dims
        if not dout.is_contiguous():
            dout = dout.contiguous()
        dx = torch.empty(B, S, C, dtype=xdt, device=dout.device)
        _call("oct_mean_pool_bwd", _p(dout), _dt(dout), _p(dx), _dt(dx), B, S, C, row0, row1, _stream())
        return dx, None, None, None


class PatchEmbedFn(torch.autograd.Function):
    """Standalone PatchEmbed.forward (vv:74-83) -> [B, T'*h*w, E] with a dense backward (wgrad over all tokens)."""

    @staticmethod
    def forward(ctx, imgs, weight, bias, p, u, act_dtype):
        _chk(imgs, weight, bias)
        B = imgs.shape[0]
        E = weight.shape[0]
        w2d = weight.view(E, -1)
        if act_dtype == torch.bfloat16:
            x = patch_embed_tc(imgs, w2d, bias, p, u, torch.bfloat16)
        else:
            patches = patchify(imgs, p, u, torch.float32)
            L = patches.shape[1]
            x = gemm(GEMM_NT, patches.view(B * L, -1), w2d, B * L, E, w2d.shape[1], torch.float32, EPI_BIAS, bias=bias)
            x = x.view(B, L, E)
        ctx.save_for_backward(imgs)
        ctx.cfg = (p, u, act_dtype, tuple(weight.shape))
        return x

    @staticmethod
    def backward(ctx, dx):
        (imgs,) = ctx.saved_tensors
        p, u, act_dtype, wshape = ctx.cfg
        if not dx.is_contiguous():
            dx = dx.contiguous()
        B, L, E = dx.shape
        patches = patchify(imgs, p, u, act_dtype).view(B * L, -1)
        dx2 = dx.view(B * L, E)
        dw, db = wgrad_bias(dx2, patches)
        return None, dw.view(wshape), db, None, None, None


def ell_from_dense(mat: torch.Tensor):
    """Dense [R, Cin] -> ELL (idx int32 [R, K], w f32 [R, K]) with K = the largest number of non-zeros in a row; column order
    ascending (a fixed summation order), padding entries have weight 0."""
    R = mat.shape[0]
    nz = mat != 0
    K = max(1, int(nz.sum(1).max()))
    idx = torch.zeros(R, K, dtype=torch.int32)
    w = torch.zeros(R, K, dtype=torch.float32)
    for r in range(R):
        cols = nz[r].nonzero().flatten()
        idx[r, :cols.numel()] = cols.to(torch.int32)
        w[r, :cols.numel()] = mat[r, cols]
    return idx.contiguous(), w.contiguous()


def ell_spmm(idx, w, x2d):
    _chk(idx, w, x2d)
    R, K = idx.shape
    C = x2d.shape[1]
    y = torch.empty(R, C, dtype=torch.float32, device=x2d.device)
    _call("oct_ell_spmm", _p(idx), _p(w), _p(x2d), _p(y), R, K, C, _stream())
    return y


class InterpTableFn(torch.autograd.Function):
    """Bicubic 32x32 -> 16x16 resampling of a learnable spatial pos-embed table (models:419-421, :537-539) as the fixed
    linear map it is:  table_lo [G_lo, C] = M [G_lo, G_hi] · table_hi [G_hi, C], with M = F.interpolate applied to the
    identity once at construction and stored sparse (16 taps per row; its transpose for the backward: <= 9):
    oct_ell_spmm, fp32, ~5 us per call instead of a 52 us dense CUDA-core GEMM.  `ell` = (idx, w, idx_T, w_T)."""

    @staticmethod
    def forward(ctx, table_hi, idx, w, idx_t, w_t):
        C = table_hi.shape[-1]
        t2 = table_hi.reshape(-1, C)
        if not t2.is_contiguous():
            t2 = t2.contiguous()
        out = ell_spmm(idx, w, t2)
        ctx.save_for_backward(idx_t, w_t)
        ctx.shape = table_hi.shape
        return out

    @staticmethod
    def backward(ctx, dout):
        idx_t, w_t = ctx.saved_tensors
        if not dout.is_contiguous():
            dout = dout.contiguous()
        d = ell_spmm(idx_t, w_t, dout.float())
        return d.view(ctx.shape), None, None, None, None
