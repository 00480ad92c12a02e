"""OCTCube-IR contrastive step: the operator surface of retinal-COEM/src/open_clip —
    CustomTextCLIP.forward / encode_image / encode_text      model.py:637-683  (two towers -> F.normalize -> logit_scale.exp())
    ClipLoss(local_loss, gather_with_grad, ...).forward      loss.py:148-229   (feature all-gather + two cross-entropies)
    gather_features                                          loss.py:21-63     (autograd-aware all_gather)
on hand-written kernels (csrc/clip.cu): the L2 normalisation, and ONE fused exchange + logits + cross-entropy kernel pair in
which the all-gather and the reduce-scatter of its backward are peer-memory loads over NVLink inside the kernels — no
collective library call, no gathered feature matrix, no [B, B*W] logits tensor.

Supported configuration = the reference recipe (src/scripts/retclip_train/train_IR_512-MAE3D-nodrop-vit-large.sh):
`--local-loss --gather-with-grad`, labels = arange, no horovod, correct_label = 0; and any world size incl. 1.  The other
ClipLoss flag combinations raise NotImplementedError.  No CPU fallback: features must be CUDA tensors.
"""
from __future__ import annotations

import ctypes
from typing import List, Optional

import torch
import torch.distributed as dist
import torch.nn as nn

from . import _lib
from .ops import _call, _chk, _dt, _p, _stream


# ----------------------------------------------------------------------------------------------------------------
# F.normalize(features, dim=-1)                                                         model.py:663,667
# ----------------------------------------------------------------------------------------------------------------
class NormalizeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, eps):
        _chk(x)
        B, D = x.shape
        y = torch.empty(B, D, dtype=torch.float32, device=x.device)
        inv = torch.empty(B, dtype=torch.float32, device=x.device)
        _call("oct_l2norm_fwd", _p(x), _dt(x), _p(y), _p(inv), B, D, float(eps), _stream())
        ctx.save_for_backward(y, inv)
        ctx.xdtype = x.dtype
        return y

    @staticmethod
    def backward(ctx, dy):
        y, inv = ctx.saved_tensors
        dy = dy.float().contiguous()
        dx = torch.empty(y.shape, dtype=ctx.xdtype, device=y.device)
        _call("oct_l2norm_bwd", _p(dy), _p(y), _p(inv), _p(dx), _dt(dx), y.shape[0], y.shape[1], _stream())
        return dx, None


def l2_normalize(x: torch.Tensor, eps: float = 1e-12) -> torch.Tensor:
    """F.normalize(x, dim=-1) for a [B, D] feature matrix (bf16 or fp32 in, fp32 out — what autocast produces)."""
    return NormalizeFn.apply(x.contiguous(), eps)


normalize = l2_normalize


# ----------------------------------------------------------------------------------------------------------------
# peer-mapped exchange buffers
# ----------------------------------------------------------------------------------------------------------------
class PeerExchange:
    """One zero-initialised device buffer per rank, mapped into every rank of `group` (one process per GPU, one node):
    CUDA IPC handles travel through the process group's object all-gather ONCE at construction; afterwards `ptrs[s]` is rank
    s's buffer as THIS process addresses it and NVLink / NVSwitch carries the loads and stores."""

    def __init__(self, nbytes: int, device, group=None):
        lib = _lib.load()
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.device = torch.device(device)
        self._opened: List[int] = []
        with torch.cuda.device(self.device):
            mine = ctypes.c_void_p()
            _lib.check(lib.oct_peer_alloc(ctypes.byref(mine), int(nbytes)), "oct_peer_alloc")
            self._mine = mine.value
            self.ptrs = [0] * self.world
            self.ptrs[self.rank] = self._mine
            if self.world > 1:
                handle = ctypes.create_string_buffer(64)
                _lib.check(lib.oct_peer_export(ctypes.c_void_p(self._mine), handle), "oct_peer_export")
                handles = [None] * self.world
                dist.all_gather_object(handles, bytes(handle.raw), group=group)
                for s, h in enumerate(handles):
                    if s == self.rank:
                        continue
                    out = ctypes.c_void_p()
                    _lib.check(lib.oct_peer_open(ctypes.create_string_buffer(h, 64), ctypes.byref(out)), "oct_peer_open")
                    self.ptrs[s] = out.value
                    self._opened.append(out.value)
                dist.barrier(group=group)  # nobody starts a step before every rank has mapped every buffer
        self.table = (ctypes.c_void_p * self.world)(*self.ptrs)

    def close(self):
        lib = _lib.load()
        if self._mine is None:
            return
        with torch.cuda.device(self.device):
            torch.cuda.synchronize()
            if self.world > 1:
                dist.barrier(group=self.group)  # peers may still be reading this rank's buffer
            for p in self._opened:
                lib.oct_peer_close(ctypes.c_void_p(p))
            lib.oct_peer_free(ctypes.c_void_p(self._mine))
        self._opened, self._mine = [], None


# ----------------------------------------------------------------------------------------------------------------
# ClipLoss                                                                                loss.py:148-229
# ----------------------------------------------------------------------------------------------------------------
class _ClipLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, image, enface, logit_scale, owner):
        _chk(image, enface, logit_scale)
        B, D = image.shape
        xch, state = owner._exchange_for(B, D, image.device)
        loss = torch.empty((), dtype=torch.float32, device=image.device)
        _call("oct_clip_loss_fwd", _p(image), _p(enface), _p(logit_scale), xch.table, _p(state), _p(loss), xch.rank, xch.world, B, D,
              _stream())
        ctx.save_for_backward(image, enface, logit_scale)
        ctx.owner = owner
        return loss

    @staticmethod
    def backward(ctx, dloss):
        image, enface, logit_scale = ctx.saved_tensors
        B, D = image.shape
        xch, state = ctx.owner._exchange_for(B, D, image.device)
        dloss = dloss.float().contiguous()
        d_image, d_enface = torch.empty_like(image), torch.empty_like(enface)
        d_scale = torch.empty_like(logit_scale)
        _call("oct_clip_loss_bwd", _p(image), _p(enface), _p(logit_scale), _p(dloss), xch.table, _p(state), _p(d_image), _p(d_enface),
              _p(d_scale), xch.rank, xch.world, B, D, _stream())
        return d_image, d_enface, d_scale, None


class ClipLoss(nn.Module):
    """Same constructor and call as the reference's ClipLoss (loss.py:150-170,178).  `rank` / `world_size` are taken from the
    process group when one is initialised (the reference passes args.rank / args.world_size, which are the same numbers)."""

    def __init__(self, local_loss=False, gather_with_grad=False, cache_labels=False, rank=0, world_size=1, use_horovod=False,
                 correct_label=0, process_group=None):
        super().__init__()
        if use_horovod or correct_label:
            raise NotImplementedError("octcubem_b200.ClipLoss: horovod / correct_label are not part of the recipe and not implemented")
        if world_size > 1 and not (local_loss and gather_with_grad):
            raise NotImplementedError("octcubem_b200.ClipLoss implements the recipe's --local-loss --gather-with-grad configuration")
        if world_size > 1 and not dist.is_initialized():
            raise RuntimeError("ClipLoss(world_size > 1) needs an initialised process group (the IPC handles travel through it)")
        if dist.is_initialized() and world_size > 1:
            assert world_size == dist.get_world_size(process_group) and rank == dist.get_rank(process_group), \
                "rank / world_size disagree with the process group"
        self.local_loss, self.gather_with_grad, self.cache_labels = local_loss, gather_with_grad, cache_labels
        self.rank, self.world_size, self.group = rank, world_size, process_group
        self._key = None
        self._xch: Optional[PeerExchange] = None
        self._state = None

    def _exchange_for(self, B, D, device):
        """Exchange buffer + device state for one (B, D): created on first use (a collective set-up step: every rank must
        reach it together, as they reach the reference's all_gather) and kept, so the step itself allocates nothing."""
        key = (int(B), int(D), device)
        if self._key != key:
            if self._xch is not None:
                self._xch.close()
            lib = _lib.load()
            self._xch = PeerExchange(lib.oct_clip_xchg_bytes(B, D), device, self.group if self.world_size > 1 else None) \
                if self.world_size > 1 else _LocalExchange(lib.oct_clip_xchg_bytes(B, D), device)
            self._state = torch.zeros(lib.oct_clip_state_bytes(B) // 4, dtype=torch.int32, device=device)
            self._key = key
        return self._xch, self._state

    def forward(self, image_features, enface_features, logit_scale):
        if image_features.shape != enface_features.shape or image_features.dim() != 2:
            raise ValueError("ClipLoss: image_features and enface_features must both be [B, D]")
        if not torch.is_tensor(logit_scale):
            logit_scale = torch.tensor(float(logit_scale), device=image_features.device)
        return _ClipLossFn.apply(image_features.float().contiguous(), enface_features.float().contiguous(),
                                 logit_scale.float().reshape(()).contiguous(), self)

    def peer_timeout(self) -> bool:
        """True if a kernel of this loss ever gave up waiting for a peer's features (synchronises; diagnostics / tests)."""
        return self._state is not None and int(self._state[3]) != 0

    def close(self):
        if self._xch is not None:
            self._xch.close()
            self._xch, self._key = None, None


class _LocalExchange:
    """world_size == 1: the 'peer' buffer is an ordinary device tensor."""

    def __init__(self, nbytes, device):
        self.buf = torch.zeros((int(nbytes) + 3) // 4, dtype=torch.int32, device=device)
        self.rank, self.world = 0, 1
        self.table = (ctypes.c_void_p * 1)(self.buf.data_ptr())

    def close(self):
        self.buf = None


# ----------------------------------------------------------------------------------------------------------------
# CustomTextCLIP                                                                          model.py:637-683
# ----------------------------------------------------------------------------------------------------------------
class CustomTextCLIP(nn.Module):
    """Two towers + temperature with the reference's call surface.  The towers are handed in already built (the reference
    builds them from a json model config, model.py:125-437 — configuration plumbing that is out of scope); for the OCTCube-IR
    recipe `visual` is octcubem_b200.models_vit_st_flash_attn.VisionTransformer with num_classes = embed_dim (its `head` is the
    512-d projection, model.py:246-270) and `text` the en-face / IR tower of the same family."""

    def __init__(self, visual: nn.Module, text: nn.Module, init_logit_scale: float = 2.6592600369327779):  # log(1 / 0.07)
        super().__init__()
        self.visual, self.text = visual, text
        self.logit_scale = nn.Parameter(torch.ones([]) * init_logit_scale)

    def encode_image(self, image, normalize: bool = False):
        features = self.visual(image)
        return l2_normalize(features) if normalize else features

    def encode_text(self, text, normalize: bool = False):
        features = self.text(text)
        return l2_normalize(features) if normalize else features

    def forward(self, image, text, single_modality=None):
        if single_modality is not None:
            assert single_modality in ["image", "text"], f"single_modality should be either 'image' or 'text', got {single_modality}"
            if single_modality == "image":
                return self.encode_image(image, normalize=True), None, self.logit_scale.exp()
            return None, self.encode_text(text, normalize=True), self.logit_scale.exp()
        return self.encode_image(image, normalize=True), self.encode_text(text, normalize=True), self.logit_scale.exp()
