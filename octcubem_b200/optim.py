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


class CosineSchedule:
    """The reference's per-iteration learning-rate rule (lr_sched.py:10-28) as data: `adjust_learning_rate(optimizer,
    data_iter_step / len(data_loader) + epoch, args)` is called at the first iteration of every accumulation group
    (engine_pretrain.py:87-91), i.e. optimizer step k (1-based) runs at the fractional epoch (k - 1) * epochs_per_step with
    epochs_per_step = accum_iter / len(data_loader).  Handed to FusedAdamW(schedule=...) the rule is evaluated ON THE DEVICE by
    oct_adamw_clock_advance, so a captured CUDA graph of the training step keeps following it when replayed."""

    def __init__(self, lr, min_lr, warmup_epochs, epochs, epochs_per_step):
        assert epochs > warmup_epochs >= 0 and epochs_per_step >= 0
        self.lr, self.min_lr, self.warmup_epochs, self.epochs = float(lr), float(min_lr), float(warmup_epochs), float(epochs)
        self.epochs_per_step = float(epochs_per_step)

    def lr_at_step(self, step: int) -> float:
        """Host-side value of the same rule (logging / tests); step counts from 1."""
        e = (step - 1) * self.epochs_per_step
        if e < self.warmup_epochs:
            return self.lr * e / self.warmup_epochs
        return self.min_lr + (self.lr - self.min_lr) * 0.5 * (
            1.0 + math.cos(math.pi * (e - self.warmup_epochs) / (self.epochs - self.warmup_epochs)))


class FusedAdamW(torch.optim.Optimizer):
    """torch.optim.AdamW semantics (decoupled weight decay, bias correction, no amsgrad / maximize), one launch per group.

    shadows: optional callable  param -> bf16 tensor of the same shape (or None); the kernel then also writes the bf16 copy
    of the updated weight (MaskedAutoencoderViT keeps such shadows for its tensor-core GEMMs)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, shadows=None, schedule=None):
        """schedule: optional CosineSchedule.  With it the step count, the bias corrections and the learning rate live in a
        16-byte device clock (`self.clock`: int32 step, then lr / bias corrections as fp32) advanced inside step(): the
        launches of one step() are then replayable from a CUDA graph (a group's `lr` is ignored, its optional `lr_scale`
        multiplies the scheduled rate; all groups must share betas).  Without it `lr` and the step count are host values
        baked into the launch — correct eagerly, frozen under graph replay."""
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._shadows = shadows
        self.schedule = schedule
        self.clock = None
        self._tables = {}
        self._norm_ws = None
        self.grad_norm = None

    def _table(self, gi, group):
        live = [p for p in group["params"] if p.grad is not None]
        shs = [self._shadows(p) if self._shadows is not None else None for p in live]
        def mom(p):  # load_state_dict() swaps the moment tensors: the chunk table must follow them
            st = self.state.get(p)
            return (st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr()) if st and "exp_avg" in st else (0, 0)
        key = tuple((p.data_ptr(), p.grad.data_ptr(), 0 if sh is None else sh.data_ptr()) + mom(p) for p, sh in zip(live, shs))
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
            for t in (p, p.grad, st["exp_avg"], st["exp_avg_sq"]):
                if t.data_ptr() % 16:
                    raise RuntimeError("FusedAdamW: parameters, gradients and moments must be 16-byte aligned (float4 access)")
            if sh is not None and sh.data_ptr() % 8:
                raise RuntimeError("FusedAdamW: bf16 shadows must be 8-byte aligned")
            n = p.numel()
            for off in range(0, n, CHUNK):
                rows.append((p.data_ptr() + 4 * off, p.grad.data_ptr() + 4 * off, st["exp_avg"].data_ptr() + 4 * off,
                             st["exp_avg_sq"].data_ptr() + 4 * off, 0 if sh is None else sh.data_ptr() + 2 * off,
                             min(CHUNK, n - off)))
        dev = group["params"][0].device
        t = torch.tensor(rows, dtype=torch.int64).reshape(-1, 6).to(dev) if rows else None
        # (the key was computed before the moments of a first step existed: store the one that matches the table)
        key = tuple((p.data_ptr(), p.grad.data_ptr(), 0 if sh is None else sh.data_ptr()) + mom(p) for p, sh in zip(live, shs))
        ent = (key, t, len(rows))
        self._tables[gi] = ent
        return ent

    def _norm_workspace(self):
        tabs = [self._table(gi, g) for gi, g in enumerate(self.param_groups) if g["params"]]
        tabs = [t for t in tabs if t[2] > 0]
        if not tabs:
            return None
        key = tuple(t[1].data_ptr() for t in tabs)
        if self._norm_ws is None or self._norm_ws[0] != key:
            allt = torch.cat([t[1] for t in tabs], 0)
            self._norm_ws = (key, allt, torch.empty(allt.shape[0], dtype=torch.float32, device=allt.device),
                             torch.empty(2, dtype=torch.float32, device=allt.device))
        return self._norm_ws

    @torch.no_grad()
    def prepare(self, max_grad_norm=None):
        """Everything step() would otherwise create on first use — the moment buffers, the chunk tables (a host-to-device
        copy), the gradient-norm workspace and the device clock — built now, without updating anything.  Call it once the
        gradients sit at their final addresses and BEFORE capturing a step into a CUDA graph: allocations and zero-fills
        recorded into the graph would be replayed (and the moments reset) on every step."""
        for gi, g in enumerate(self.param_groups):
            if g["params"]:
                self._table(gi, g)
        if max_grad_norm is not None:
            self._norm_workspace()
        if self.schedule is not None and self.clock is None:
            dev = next(p for g in self.param_groups for p in g["params"]).device
            self.clock = torch.zeros(4, dtype=torch.int32, device=dev)
        return self

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
            ws = self._norm_workspace()
            if ws is not None:
                _, allt, partial, out = ws
                rc = lib.oct_grad_norm(ctypes.c_void_p(allt.data_ptr()), allt.shape[0], float(grad_scale), float(max_grad_norm),
                                       ctypes.c_void_p(partial.data_ptr()), ctypes.c_void_p(out.data_ptr()),
                                       ctypes.c_void_p(torch.cuda.current_stream(allt.device).cuda_stream))
                if rc != 0:
                    raise RuntimeError(f"oct_grad_norm failed (code {rc}): {_lib.last_error()}")
                self.grad_norm = out[0]
                clip_ptr = ctypes.c_void_p(out[1:].data_ptr())
        if self.schedule is not None:
            self._step_clocked(lib, grad_scale, clip_ptr)
            return loss
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

    # ------------------------------------------------------------------ graph-replayable path (device clock)
    def _step_clocked(self, lib, grad_scale, clip_ptr):
        groups = [(gi, g) for gi, g in enumerate(self.param_groups) if g["params"]]
        if not groups:
            return
        dev = groups[0][1]["params"][0].device
        betas = {tuple(g["betas"]) for _, g in groups}
        if len(betas) != 1:
            raise RuntimeError("FusedAdamW(schedule=...): all parameter groups must share betas (one device clock)")
        b1, b2 = betas.pop()
        if self.clock is None:
            self.clock = torch.zeros(4, dtype=torch.int32, device=dev)
        sc = self.schedule
        st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        rc = lib.oct_adamw_clock_advance(ctypes.c_void_p(self.clock.data_ptr()), sc.lr, sc.min_lr, sc.warmup_epochs, sc.epochs,
                                         sc.epochs_per_step, float(b1), float(b2), st)
        if rc != 0:
            raise RuntimeError(f"oct_adamw_clock_advance failed (code {rc}): {_lib.last_error()}")
        for gi, group in groups:
            _, table, n = self._table(gi, group)
            if n == 0:
                continue
            rc = lib.oct_adamw_step_clocked(ctypes.c_void_p(table.data_ptr()), n, ctypes.c_void_p(self.clock.data_ptr()),
                                            float(group.get("lr_scale", 1.0)), float(b1), float(b2), float(group["eps"]),
                                            float(group["weight_decay"]), float(grad_scale), clip_ptr, st)
            if rc != 0:
                raise RuntimeError(f"oct_adamw_step_clocked failed (code {rc}): {_lib.last_error()}")
            touched = [p for p in group["params"] if p.grad is not None]
            for p in touched:
                self.state[p]["step"] += 1  # host mirror: exact for eager steps, the device clock is authoritative under replay
            torch._C._autograd._unsafe_set_version_counter(touched, [p._version + 1 for p in touched])

    # ------------------------------------------------------------------ checkpoint / resume (misc.save_model / load_model)
    def state_dict(self):
        """torch.optim.Optimizer.state_dict() with the step count read back from the DEVICE clock: under CUDA-graph replay the
        host mirror `state[p]['step']` does not advance, the clock is authoritative (one host sync, checkpoint time only)."""
        if self.schedule is not None and self.clock is not None:
            step, _ = self.clock_state()
            for group in self.param_groups:
                for p in group["params"]:
                    st = self.state.get(p)
                    if st and "step" in st:
                        st["step"] = step
        return super().state_dict()

    def load_state_dict(self, state_dict):
        """Restores moments and step; the chunk tables / norm workspace are rebuilt on the next step (they hold the OLD moment
        pointers) and the device clock restarts from the loaded step, so the cosine schedule and the bias corrections resume
        where the checkpoint left them."""
        super().load_state_dict(state_dict)
        self._tables, self._norm_ws = {}, None
        steps = {int(st["step"]) for st in self.state.values() if "step" in st}
        if len(steps) > 1:
            raise RuntimeError(f"FusedAdamW.load_state_dict: parameters disagree on the step count ({sorted(steps)})")
        for st in self.state.values():
            if "step" in st:
                st["step"] = int(st["step"])            # torch's AdamW stores 0-d tensors; this optimizer counts in python ints
        if self.schedule is not None and steps:
            self.set_clock(steps.pop())

    def set_clock(self, step: int):
        """Positions the device clock: the NEXT step() is optimizer step `step + 1` (lr and bias corrections follow from it)."""
        dev = next(p for g in self.param_groups for p in g["params"]).device
        if self.clock is None:
            self.clock = torch.zeros(4, dtype=torch.int32, device=dev)
        self.clock.copy_(torch.tensor([int(step), 0, 0, 0], dtype=torch.int32), non_blocking=False)

    def writes_shadows(self) -> bool:
        """True when step() also emits the bf16 weight shadows (constructed with `shadows=`)."""
        return self._shadows is not None

    def clock_state(self):
        """(step, lr) read back from the device clock (synchronises; logging only, engine_pretrain.py:186-187)."""
        if self.clock is None:
            return 0, 0.0
        raw = self.clock.cpu()
        return int(raw[0]), float(raw[1:2].view(torch.float32)[0])
