"""The optimizer step of the reference's training loop as ONE replayable CUDA graph.

Mirrors the loop body of `train_one_epoch_joint` (Pre-training/engine_pretrain.py:83-161) — the caller on the input side of
the hot path:
    lr_sched.adjust_learning_rate(...)                              :87-91   -> device clock of optim.FusedAdamW
    samples.reshape(b * r, c, t, h, w)                              :101-103
    feat = forward_patch_embed(samples); misc.get_mask(feat)        :112-115 -> dropped: `forward` discards pre_mask
                                                                                (models...:677, quirk Q7), the pass is dead
    loss, _, _ = model(samples, mask_ratio, frame_loss=True, ...)   :117-122
    loss_2d, _, _ = model(sample_2d, mask_ratio=mask_ratio_2d)      :124-127
    loss = loss + loss_2d; loss /= accum_iter                       :148-163
    loss_scaler(loss, optimizer, clip_grad=..., update_grad=...)    :164-170 -> backward, grad-norm clip, AdamW
    optimizer.zero_grad()                                           :172-173
    loss.item() x3, frame_loss.item() per frame, cuda.synchronize() :130-147,175 -> no host sync inside the step: results
                                                                                stay on the device until the caller reads them
The first `warm_steps` calls with a new input signature run eagerly (they are real training steps: lazy initialisation,
the reducer's bucket discovery); the next call captures the whole step — both forwards, backward, the gradient all-reduce,
the device-side grad-norm clip and the AdamW update — and from then on the step is one graph launch.  Input signatures
(volume shape, 2D batch shape, the 2D keep count that `mask_ratio_2d_scheduler` changes from epoch to epoch) each get their
own graph.  bf16 needs no loss scaling, so NativeScaler's GradScaler is not reproduced; its clip_grad / grad-norm output is.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, Optional

import torch

from . import ops
from .dp import GradReducer
from .models_mae import len_keep_of
from .optim import FusedAdamW


def K_scheduler(epoch, K_max=0.7, K_min=0.3, all_epoch=100, warmup_epochs=10, epoch_offset=0):
    """main_pretrain_oph_joint_2d512_flash_attn.py:53-59: starts at K_max, decreases linearly to K_min after the warm-up."""
    n = epoch - epoch_offset
    if n <= warmup_epochs:
        return K_max
    return K_max - (n - warmup_epochs) * (K_max - K_min) / (all_epoch - warmup_epochs - epoch_offset)


def mask_ratio_2d_scheduler(epoch, mask_ratio_max=0.85, mask_ratio_min=0.75, all_epoch=100, warmup_epochs=10, epoch_offset=0):
    """main_pretrain_oph_joint_2d512_flash_attn.py:61-67: starts at mask_ratio_min, increases linearly to mask_ratio_max."""
    n = epoch - epoch_offset
    if n <= warmup_epochs:
        return mask_ratio_min
    return mask_ratio_min + (n - warmup_epochs) * (mask_ratio_max - mask_ratio_min) / (all_epoch - warmup_epochs - epoch_offset)


class StepResult:
    """Device-resident results of one step (nothing has been copied to the host yet)."""

    def __init__(self, loss, loss_2d, loss_all, frame_loss, grad_norm):
        self.loss, self.loss_2d, self.loss_all, self.frame_loss, self.grad_norm = loss, loss_2d, loss_all, frame_loss, grad_norm

    def check_finite(self) -> Dict[str, float]:
        """The reference's per-iteration host read (engine_pretrain.py:130-161): returns the python floats and raises like
        `raise Exception("Loss is {}, stopping training")` when the 3D loss is not finite.  Synchronises."""
        out = {"loss": float(self.loss), "loss_all": float(self.loss_all)}
        if self.loss_2d is not None:
            out["loss_2d"] = float(self.loss_2d)
        if self.grad_norm is not None:
            out["grad_norm"] = float(self.grad_norm)
        if not math.isfinite(out["loss"]):
            raise Exception("Loss is {}, stopping training".format(out["loss"]))
        return out


class JointPretrainStep:
    def __init__(self, model, optimizer: FusedAdamW, mask_ratio: float = 0.9, clip_grad: Optional[float] = None,
                 accum_iter: int = 1, use_graph: bool = True, warm_steps: int = 2, process_group=None, max_graphs: int = 2):
        if int(accum_iter) < 1:
            raise ValueError("accum_iter must be >= 1")
        if use_graph and optimizer.schedule is None:
            raise ValueError("JointPretrainStep(use_graph=True) needs FusedAdamW(schedule=CosineSchedule(...)): a host-side "
                             "learning rate / step count would be frozen into the captured graph")
        self.model, self.optimizer = model, optimizer
        self.mask_ratio, self.clip_grad = mask_ratio, clip_grad
        # gradient accumulation (engine_pretrain.py:163-173): `loss /= accum_iter`, the optimizer (and, here, the gradient
        # exchange) runs on every accum_iter-th call only, gradients are zeroed after it.  Calls are counted per object.
        self.accum_iter, self._micro = int(accum_iter), 0
        self.use_graph, self.warm_steps = use_graph, max(2, int(warm_steps))  # bucket discovery + one bucketed step
        self.reducer = GradReducer(model, process_group)
        # graphs per input signature, least recently used first.  `mask_ratio_2d_scheduler` walks through ~100 keep counts
        # over a run (one per epoch): only the newest `max_graphs` signatures keep their graph, and all graphs allocate
        # from ONE private pool (they replay strictly one after another), so memory does not grow with the epoch count
        self.max_graphs = max(1, int(max_graphs))
        self._entries: "OrderedDict[tuple, dict]" = OrderedDict()
        self._pool = None
        self._joint: Optional[bool] = None
        # the bf16 path's GEMMs read bf16 weight shadows: either the optimizer writes them (shadows=model.shadow_of) and the
        # model is told so after every step, or nobody marks them current and the forward re-casts from the fp32 masters
        self._opt_writes_shadows = bool(getattr(optimizer, "writes_shadows", lambda: False)())

    # ------------------------------------------------------------------ one step, eager or under capture
    def _run(self, vol, img, ratio_2d, noise, noise_2d, out, first=True, last=True):
        if first:
            self.reducer.zero_grad()
        (loss, frame_loss), _, _ = self.model(vol, mask_ratio=self.mask_ratio, frame_loss=True, noise=noise)
        total = loss
        if img is not None:
            loss_2d, _, _ = self.model(img, mask_ratio=ratio_2d, noise=noise_2d)
            total = loss + loss_2d
            out["loss_2d"].copy_(loss_2d.detach())
        if last:
            self.reducer.backward(total, scale=1.0 / self.accum_iter)
            self.reducer.finish()
            self.optimizer.step(max_grad_norm=self.clip_grad)
            if self._opt_writes_shadows:
                self.model.shadows_current()
        else:
            with self.reducer.no_sync():                       # a micro-step: the sums stay local, nothing is exchanged
                self.reducer.backward(total, scale=1.0 / self.accum_iter)
        out["loss"].copy_(loss.detach())
        out["loss_all"].copy_(total.detach())
        out["frame_loss"].copy_(frame_loss)
        if self.clip_grad is not None and self.optimizer.grad_norm is not None:
            out["grad_norm"].copy_(self.optimizer.grad_norm)

    def _result(self, out, joint):
        return StepResult(out["loss"], out["loss_2d"] if joint else None, out["loss_all"], out["frame_loss"],
                          out["grad_norm"] if self.clip_grad is not None else None)

    def __call__(self, samples, sample_2d=None, mask_ratio_2d: float = 0.75, noise=None, noise_2d=None, flips=None) -> StepResult:
        """samples [b,c,t,h,w] or [b,r,c,t,h,w] (engine_pretrain.py:101-103), sample_2d [b2,1,3,H,W] or None, both already on
        the device.  `noise` / `noise_2d` (optional) replace the models' torch.rand draws for reproducible masks.
        A uint8 `samples` [b, T_src, H, W] is the raw cube (1 byte per pixel over PCIe): ops.ingest_u8 scales, centre-pads /
        crops it to the model's frame count and applies the per-sample `flips = (flip_t, flip_w)` ([b] uint8 flags or None)
        while writing the step's fp32 input — on replayed steps straight into the graph's static buffer."""
        cube = None
        if samples.dtype == torch.uint8:
            cube, frames = samples, self.model.patch_embed.frames
            flip_t, flip_w = flips if flips is not None else (None, None)
            samples = torch.empty(0)  # placeholder: only the shape below is used until the cube is ingested
            vol_shape = (cube.shape[0], 1, frames, cube.shape[2], cube.shape[3])
        elif samples.dim() == 6:
            b, r, c, t, h, w = samples.shape
            samples = samples.reshape(b * r, c, t, h, w)
        if cube is None:
            vol_shape = tuple(samples.shape)
        joint = sample_2d is not None
        if self._joint is not None and joint != self._joint:
            # another set of parameters takes part (quirk Q13): rebuild the buckets, forget graphs that point into the old ones
            self.reducer.reset()
            self._entries.clear()
        self._joint = joint
        dev = cube.device if cube is not None else samples.device
        pe = self.model.high_res_patch_embed if joint else None
        keep_2d = len_keep_of(pe.input_size[1] * pe.input_size[2], mask_ratio_2d) if joint else None
        first, last = self._micro == 0, self._micro == self.accum_iter - 1
        self._micro = (self._micro + 1) % self.accum_iter
        key = (vol_shape, tuple(sample_2d.shape) if joint else None, keep_2d, noise is not None, noise_2d is not None, first, last)
        ent = self._entries.get(key)
        if ent is not None:
            self._entries.move_to_end(key)
        if ent is None:
            tp = vol_shape[2] // self.model.patch_embed.t_patch_size
            out = {"loss": torch.zeros((), device=dev), "loss_2d": torch.zeros((), device=dev),
                   "loss_all": torch.zeros((), device=dev), "frame_loss": torch.zeros(vol_shape[0], tp, device=dev),
                   "grad_norm": torch.zeros((), device=dev)}
            ent = self._entries[key] = {"calls": 0, "graph": None, "out": out}
            while len(self._entries) > self.max_graphs * min(self.accum_iter, 3):   # (one graph per micro-step phase)
                _, old = self._entries.popitem(last=False)
                old.clear()
        out = ent["out"]
        if not self.use_graph or ent["calls"] < self.warm_steps:
            ent["calls"] += 1
            if cube is not None:
                samples = ops.ingest_u8(cube, vol_shape[2], flip_t=flip_t, flip_w=flip_w)
            self._run(samples, sample_2d, mask_ratio_2d, noise, noise_2d, out, first, last)
            return self._result(out, joint)
        if ent["graph"] is None:
            if cube is not None:
                samples = ops.ingest_u8(cube, vol_shape[2], flip_t=flip_t, flip_w=flip_w)
            st = ent["static"] = {"vol": samples.clone(), "img": sample_2d.clone() if joint else None,
                                  "noise": noise.clone() if noise is not None else None,
                                  "noise_2d": noise_2d.clone() if noise_2d is not None else None}
            if last:
                self.optimizer.prepare(max_grad_norm=self.clip_grad)
            torch.cuda.synchronize(dev)
            graph = torch.cuda.CUDAGraph()
            if self._pool is None:
                self._pool = torch.cuda.graph_pool_handle()
            with torch.cuda.graph(graph, pool=self._pool):  # capture records, it does not execute
                self._run(st["vol"], st["img"], mask_ratio_2d, st["noise"], st["noise_2d"], out, first, last)
            ent["graph"] = graph
        else:
            st = ent["static"]
            if cube is not None:
                ops.ingest_u8(cube, vol_shape[2], out=st["vol"], flip_t=flip_t, flip_w=flip_w)
            else:
                st["vol"].copy_(samples, non_blocking=True)
            if joint:
                st["img"].copy_(sample_2d, non_blocking=True)
            if noise is not None:
                st["noise"].copy_(noise, non_blocking=True)
            if noise_2d is not None:
                st["noise_2d"].copy_(noise_2d, non_blocking=True)
        ent["graph"].replay()
        return self._result(out, joint)

    def close(self):
        self._entries.clear()
        self.reducer.remove()
