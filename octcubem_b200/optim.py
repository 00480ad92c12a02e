"""Optimizer side of the pre-training step: the reference builds two parameter groups with
`misc.add_weight_decay` (custom_util/misc.py:678-696), runs `torch.optim._multi_tensor.AdamW(param_groups, lr, betas=(0.9, 0.95))`
(main_pretrain_oph_joint_2d512_flash_attn.py:442-455) and sets the learning rate per iteration with the half-cycle cosine of
custom_util/lr_sched.py:10-28.  `FusedAdamW` does the whole parameter update of a group in ONE kernel launch
(csrc/optim.cu through `oct_adamw_step`), optionally unscaling the gradient first and emitting the bf16 weight shadows the
next forward needs.  There is no CPU fallback: parameters must live on a CUDA device.
"""
from __future__ import annotations

import ctypes
import math

import torch

from . import _lib

CHUNK = 16384


def add_weight_decay(model, weight_decay=1e-5, skip_list=(), bias_wd=False):
    """Same grouping rule as the reference (misc.py:678-696): 1-D tensors (unless bias_wd), `*.bias` and names in skip_list
    get no weight decay.  Returns [{no_decay}, {decay}] in that order."""
    decay, no_decay = [], []
    for name, p in model.named_parameters():
        if not p.requires_grad:
            continue
        if ((not bias_wd) and p.dim() == 1) or name.endswith(".bias") or name in skip_list:
            no_decay.append(p)
        else:
            decay.append(p)
    return [{"params": no_decay, "weight_decay": 0.0}, {"params": decay, "weight_decay": weight_decay}]


def adjust_learning_rate(optimizer, epoch, lr, min_lr, warmup_epochs, epochs):
    """Half-cycle cosine after linear warm-up, evaluated per iteration with a fractional epoch (lr_sched.py:10-28);
    honours a per-group `lr_scale`."""
    if epoch < warmup_epochs:
        cur = lr * epoch / warmup_epochs
    else:
        cur = min_lr + (lr - min_lr) * 0.5 * (1.0 + math.cos(math.pi * (epoch - warmup_epochs) / (epochs - warmup_epochs)))
    for g in optimizer.param_groups:
        g["lr"] = cur * g["lr_scale"] if "lr_scale" in g else cur
    return cur


class FusedAdamW(torch.optim.Optimizer):
    """torch.optim.AdamW semantics (decoupled weight decay, bias correction, no amsgrad / maximize), one launch per group.

    shadows: optional callable  param -> bf16 tensor of the same shape (or None); the kernel then also writes the bf16 copy
    of the updated weight (MaskedAutoencoderViT keeps such shadows for its tensor-core GEMMs)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, shadows=None):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._shadows = shadows
        self._tables = {}
        self._norm_ws = None
        self.grad_norm = None

    def _table(self, gi, group):
        live = [p for p in group["params"] if p.grad is not None]
        shs = [self._shadows(p) if self._shadows is not None else None for p in live]
        key = tuple((p.data_ptr(), p.grad.data_ptr(), 0 if sh is None else sh.data_ptr()) for p, sh in zip(live, shs))
        ent = self._tables.get(gi)
        if ent is not None and ent[0] == key:
            return ent
        # (re)build: first step, gradients moved (plain autograd allocates new ones) or the participating set changed
        rows = []
        for p, sh in zip(live, shs):
            if not p.is_cuda or p.dtype != torch.float32 or p.grad.dtype != torch.float32:
                raise RuntimeError("FusedAdamW: fp32 CUDA parameters and gradients only (there is no CPU fallback)")
            if not (p.is_contiguous() and p.grad.is_contiguous()):
                raise RuntimeError("FusedAdamW: parameters and gradients must be contiguous")
            st = self.state[p]
            if not st:
                st["step"] = 0
                st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            n = p.numel()
            for off in range(0, n, CHUNK):
                rows.append((p.data_ptr() + 4 * off, p.grad.data_ptr() + 4 * off, st["exp_avg"].data_ptr() + 4 * off,
                             st["exp_avg_sq"].data_ptr() + 4 * off, 0 if sh is None else sh.data_ptr() + 2 * off,
                             min(CHUNK, n - off)))
        dev = group["params"][0].device
        t = torch.tensor(rows, dtype=torch.int64).reshape(-1, 6).to(dev) if rows else None
        ent = (key, t, len(rows))
        self._tables[gi] = ent
        return ent

    @torch.no_grad()
    def step(self, closure=None, grad_scale: float = 1.0, max_grad_norm=None):
        """grad_scale: multiplied into every gradient first (1 / loss_scale of a GradScaler).  max_grad_norm: clip the global
        gradient norm like torch.nn.utils.clip_grad_norm_ (misc.py:326-344); the norm is computed on the device, left in
        `self.grad_norm` (a 0-d CUDA tensor, as the reference's NativeScaler returns it) and never read by the host."""
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.load()
        from . import ops
        ops.join_wgrad()  # weight gradients are produced on a side stream (normally already joined at the end of backward)
        clip_ptr = None
        if max_grad_norm is not None:
            tabs = [self._table(gi, g) for gi, g in enumerate(self.param_groups) if g["params"]]
            tabs = [t for t in tabs if t[2] > 0]
            if tabs:
                key = tuple(t[1].data_ptr() for t in tabs)
                if self._norm_ws is None or self._norm_ws[0] != key:
                    allt = torch.cat([t[1] for t in tabs], 0)
                    self._norm_ws = (key, allt, torch.empty(allt.shape[0], dtype=torch.float32, device=allt.device),
                                     torch.empty(2, dtype=torch.float32, device=allt.device))
                _, allt, partial, out = self._norm_ws
                rc = lib.oct_grad_norm(ctypes.c_void_p(allt.data_ptr()), allt.shape[0], float(grad_scale), float(max_grad_norm),
                                       ctypes.c_void_p(partial.data_ptr()), ctypes.c_void_p(out.data_ptr()),
                                       ctypes.c_void_p(torch.cuda.current_stream(allt.device).cuda_stream))
                if rc != 0:
                    raise RuntimeError(f"oct_grad_norm failed (code {rc}): {_lib.last_error()}")
                self.grad_norm = out[0]
                clip_ptr = ctypes.c_void_p(out[1:].data_ptr())
        for gi, group in enumerate(self.param_groups):
            if not group["params"]:
                continue
            _, table, n = self._table(gi, group)
            if n == 0:
                continue
            steps = {self.state[p]["step"] for p in group["params"] if p.grad is not None}
            if len(steps) != 1:
                raise RuntimeError("FusedAdamW: parameters of one group must share their step count")
            step = steps.pop() + 1
            b1, b2 = group["betas"]
            dev = group["params"][0].device
            rc = lib.oct_adamw_step(ctypes.c_void_p(table.data_ptr()), n, float(group["lr"]), float(b1), float(b2),
                                    float(group["eps"]), float(group["weight_decay"]), step, float(grad_scale), clip_ptr,
                                    ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
            if rc != 0:
                raise RuntimeError(f"oct_adamw_step failed (code {rc}): {_lib.last_error()}")
            touched = [p for p in group["params"] if p.grad is not None]
            for p in touched:
                self.state[p]["step"] = step
            # the kernel wrote the parameters behind autograd's back: bump their version counters like an in-place op
            # would (the model's bf16 shadow cache and autograd's saved-tensor checks key on them)
            torch._C._autograd._unsafe_set_version_counter(touched, [p._version + 1 for p in touched])
        return loss
