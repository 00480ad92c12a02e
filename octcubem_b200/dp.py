"""Data-parallel gradient reducer: replaces torch.nn.parallel.DistributedDataParallel at the reference's call site
(Pre-training/main_pretrain_oph_joint_2d512_flash_attn.py:435-439, init custom_util/misc.py:252-297).

DDP semantics are kept: every rank computes its local masked-mean loss, gradients are summed over ranks
and divided by the world size.  Mechanics are B200-first rather than DDP's:
  * buckets are laid out in BACKWARD order (decoder first, patch-embed / pos tables last) so that the only exposed
    communication is the small final bucket;
  * each parameter's gradient is moved into its flat fp32 bucket by a post-accumulate hook, `param.grad` becomes a
    view of the bucket (no second copy after the collective);
  * a bucket is all-reduced on a dedicated side stream the moment its last gradient lands — NCCL over NVLink 5 /
    NVSwitch runs underneath the remaining backward kernels; `finish()` joins the streams;
  * `reducer.backward(loss)` seeds the backward pass with 1 / world, so the SUM all-reduce already yields DDP's mean and
    no scaling pass over the 1.3 GB of gradients follows the collective (a power-of-two scale: bit-identical);
  * the Linear / Mlp weight gradients — 99 % of the bytes — are written by the wgrad GEMM straight into their bucket
    views (`ops.grad_sinks`), so the hook has nothing to copy;
  * parameters that take no part in the step (the two high-res patch-embed tensors on a 3D-only step, quirk Q13) are
    discovered on the first, non-overlapped, step and left out of the buckets.
Works on CPU tensors with the gloo backend (used by the world_size-2 tests) — there the "side stream" is implicit.
"""
from __future__ import annotations

import contextlib
from typing import Dict, List, Optional

import torch
import torch.distributed as dist


class _Bucket:
    ALIGN = 32  # elements (fp32): 128 bytes

    @classmethod
    def padded_numel(cls, params):
        return sum(-(-p.numel() // cls.ALIGN) * cls.ALIGN for p in params)

    def __init__(self, params: List[torch.nn.Parameter], names: List[str], device, dtype, flat=None, arena_off=0):
        self.params, self.names = params, names
        # every view starts on a 128-byte boundary: the optimizer / grad-norm kernels use float4 accesses on `param.grad` and
        # the wgrad kernels reach the views through TMA (16-byte alignment), whatever the numel of the tensors in front
        offs, off = [], 0
        for p in params:
            offs.append(off)
            off += -(-p.numel() // self.ALIGN) * self.ALIGN
        self.numel = off
        # `flat`: a slice of the reducer's symmetric arena (element offset arena_off), else an ordinary tensor
        self.flat = flat if flat is not None else torch.zeros(self.numel, dtype=dtype, device=device)
        self.arena_off = arena_off                                        # (padding stays zero: a SUM all-reduce keeps it zero)
        self.views = [self.flat[o:o + p.numel()].view_as(p) for o, p in zip(offs, params)]
        self.pending = len(params)
        self.launched = False
        self.work = None


class GradReducer:
    def __init__(self, model: torch.nn.Module, process_group=None, bucket_mb: float = 32.0, first_bucket_mb: float = 8.0,
                 last_bucket_mb: float = 4.0):
        self.model = model
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.bucket_bytes = int(bucket_mb * 2 ** 20)
        self.first_bucket_bytes = int(first_bucket_mb * 2 ** 20)
        self.last_bucket_bytes = int(last_bucket_mb * 2 ** 20)
        self.buckets: Optional[List[_Bucket]] = None
        self._slot: Dict[int, tuple] = {}
        self._hooks = []
        self._order: List[str] = []   # order in which gradients became ready on the discovery step
        self._seen = set()
        self._named = dict(model.named_parameters())
        self._sink_keys: List[int] = []
        self._discovering = True
        self._prescaled = False
        self._stream_used = False
        self._no_sync = False
        dev = next(model.parameters()).device
        self.device = dev
        self.stream = torch.cuda.Stream(device=dev) if dev.type == "cuda" else None
        for name, p in self._named.items():
            if p.requires_grad:
                self._hooks.append(p.register_post_accumulate_grad_hook(self._make_hook(name)))

    # ------------------------------------------------------------------ bucket construction
    def _build(self, ready_order: List[str]):
        """ready_order: parameter names in the order their gradients appeared (backward order)."""
        sizes = [self._named[n].numel() * 4 for n in ready_order]
        # the tail bucket = the last-ready parameters, at most last_bucket_bytes: the only exposed communication
        tail_start, acc = len(ready_order), 0
        while tail_start > 0 and (acc + sizes[tail_start - 1] <= self.last_bucket_bytes or tail_start == len(ready_order)):
            tail_start -= 1
            acc += sizes[tail_start]
        groups, cur, cur_bytes = [], [], 0
        for i in range(tail_start):
            limit = self.first_bucket_bytes if not groups else self.bucket_bytes  # small first bucket starts the pipe early
            if cur and cur_bytes + sizes[i] > limit:
                groups.append(cur)
                cur, cur_bytes = [], 0
            cur.append(ready_order[i])
            cur_bytes += sizes[i]
        if cur:
            groups.append(cur)
        if tail_start < len(ready_order):
            groups.append(list(ready_order[tail_start:]))
        plists = [[self._named[n] for n in g] for g in groups]
        arena = self._symmetric_arena(sum(_Bucket.padded_numel(pl) for pl in plists))
        buckets, off = [], 0
        for g, pl in zip(groups, plists):
            n = _Bucket.padded_numel(pl)
            buckets.append(_Bucket(pl, g, self.device, torch.float32, None if arena is None else arena[off:off + n], off))
            off += n
        self.buckets = buckets
        self._slot = {}
        for bi, b in enumerate(buckets):
            for pi, p in enumerate(b.params):
                self._slot[id(p)] = (bi, pi)
        self._register_sinks()

    # ------------------------------------------------------------------ symmetric memory (NVLink / NVSwitch all-reduce)
    def _symmetric_arena(self, numel):
        """One symmetric allocation for ALL buckets (+ the flag block of csrc/allreduce.cu): every rank's copy is mapped into
        every process, and behind one multicast address where the fabric supports it.  Returns the local fp32 tensor (the
        buckets are slices of it) or None -> ordinary tensors + NCCL (CPU / gloo, world 1, OCT_ALLREDUCE=nccl, or a platform
        where the rendezvous fails — reported once on rank 0, never silent)."""
        import os
        self._sym = None
        if self.device.type != "cuda" or self.world == 1 or os.environ.get("OCT_ALLREDUCE", "sym") == "nccl":
            return None
        try:
            import ctypes
            import torch.distributed._symmetric_memory as symm
            from . import _lib
            lib = _lib.load()
            flag_elems = lib.oct_allreduce_flag_bytes() // 4
            total = numel + flag_elems
            arena = symm.empty(total, dtype=torch.float32, device=self.device)
            group = self.pg if self.pg is not None else dist.group.WORLD
            hdl = symm.rendezvous(arena, group)
            arena.zero_()
            torch.cuda.synchronize(self.device)
            dist.barrier(group=self.pg)                                  # flags are zero everywhere before anybody signals
            ptrs = [int(p) for p in hdl.buffer_ptrs]
            mc = 0 if os.environ.get("OCT_ALLREDUCE", "sym") == "peer" else int(getattr(hdl, "multicast_ptr", 0) or 0)
            self._sym = {"arena": arena, "hdl": hdl, "mc": mc, "table": (ctypes.c_void_p * self.world)(*ptrs),
                         "flag_off": numel * 4, "rank": dist.get_rank(self.pg),
                         "state": torch.zeros(lib.oct_allreduce_state_bytes() // 4, dtype=torch.int32, device=self.device),
                         "ctas": int(os.environ.get("OCT_AR_CTAS", "96"))}
            if dist.get_rank(self.pg) == 0 and os.environ.get("OCT_VERBOSE"):
                import sys
                print(f"[octcubem_b200] gradient all-reduce: symmetric memory, {'multicast (NVLS)' if mc else 'peer loads/stores'}, "
                      f"{total * 4 / 2**20:.0f} MiB per rank", file=sys.stderr, flush=True)
            return arena[:numel]
        except Exception as e:  # noqa: BLE001
            if dist.get_rank(self.pg) == 0:
                import sys
                print(f"[octcubem_b200] symmetric-memory all-reduce unavailable ({type(e).__name__}: {e}); using NCCL",
                      file=sys.stderr, flush=True)
            self._sym = None
            return None

    def allreduce_backend(self) -> str:
        if getattr(self, "_sym", None) is None:
            return "nccl" if self.world > 1 else "none"
        return "symmetric-multicast" if self._sym["mc"] else "symmetric-peer"

    def peer_timeout(self) -> bool:
        """True if an all-reduce kernel ever gave up waiting for a peer (synchronises; diagnostics)."""
        return getattr(self, "_sym", None) is not None and bool(int(self._sym["state"].view(-1, 4)[:, 2].sum()) != 0)

    def _register_sinks(self):
        """CUDA path: the autograd Functions in ops.py write weight / bias gradients directly into the bucket views."""
        if self.device.type != "cuda":
            return
        from . import ops
        for b in self.buckets:
            for p, v in zip(b.params, b.views):
                ops.grad_sinks[p.data_ptr()] = (v, p)
                self._sink_keys.append(p.data_ptr())

    def _clear_sinks(self):
        if self._sink_keys:
            from . import ops
            for k in self._sink_keys:
                ops.grad_sinks.pop(k, None)
        self._sink_keys = []

    def bucket_layout(self):
        """[(parameter names, payload elements)] per bucket (the flat buffers are a little longer: 128-byte aligned views)."""
        return [(b.names, sum(p.numel() for p in b.params)) for b in (self.buckets or [])]

    # ------------------------------------------------------------------ hooks
    def _make_hook(self, name):
        def hook(p):
            if self._discovering:
                if name not in self._seen:                 # (an accumulation group visits every parameter once per micro-step)
                    self._seen.add(name)
                    self._order.append(name)
                return
            slot = self._slot.get(id(p))
            if slot is None:
                raise RuntimeError(f"GradReducer: parameter {name} produced a gradient but was unused on the discovery step "
                                   "(call reset() when the set of participating parameters changes)")
            b = self.buckets[slot[0]]
            view = b.views[slot[1]]
            if p.grad.data_ptr() != view.data_ptr():
                view.copy_(p.grad)
                p.grad = view
            if self._no_sync:  # a micro-step of an accumulation group: keep the sum local
                return
            b.pending -= 1
            if b.pending == 0:
                self._launch(b)
        return hook

    def _launch(self, b: _Bucket):
        b.launched = True
        if self.world == 1:
            return
        if self.stream is not None:
            self._stream_used = True
            self.stream.wait_stream(torch.cuda.current_stream(self.device))
            from . import ops
            if ops._wgrad_pending.get(self.device.index, False):  # weight gradients are produced on their own stream
                self.stream.wait_stream(ops.wgrad_stream(self.device))
            with torch.cuda.stream(self.stream):
                sym = getattr(self, "_sym", None)
                if sym is not None:
                    import ctypes
                    from . import _lib
                    rc = _lib.load().oct_allreduce_sym(
                        ctypes.c_void_p(sym["mc"]) if sym["mc"] else None, sym["table"], sym["flag_off"],
                        ctypes.c_void_p(sym["state"].data_ptr()), b.arena_off, b.numel, sym["rank"], self.world, 0,
                        1.0 if self._prescaled else 1.0 / self.world, sym["ctas"], ctypes.c_void_p(self.stream.cuda_stream))
                    _lib.check(rc, "oct_allreduce_sym")
                else:
                    dist.all_reduce(b.flat, op=dist.ReduceOp.SUM, group=self.pg)
                    if not self._prescaled:
                        b.flat.mul_(1.0 / self.world)
        else:
            b.work = dist.all_reduce(b.flat, op=dist.ReduceOp.SUM, group=self.pg, async_op=True)

    # ------------------------------------------------------------------ step protocol
    @contextlib.contextmanager
    def no_sync(self):
        """Micro-steps 1 .. k-1 of a gradient-accumulation group (the role of DistributedDataParallel.no_sync): gradients
        accumulate in the buckets, nothing is reduced, finish() is not called.  The LAST micro-step runs outside this context
        and reduces the accumulated sums.  Use the same backward style (reducer.backward or loss.backward) for all of them.
        During the bucket-discovery group the micro-steps accumulate in ordinary autograd gradients; finish() of the group's
        last pass moves the sums into the freshly built buckets."""
        self._no_sync = True
        try:
            yield
        finally:
            self._no_sync = False

    def backward(self, loss: torch.Tensor, scale: float = 1.0):
        """loss.backward() with the gradient seeded at scale / world: the summed gradients are DDP's averages as they leave the
        collective (`scale` = 1 / accum_iter folds the reference's `loss /= accum_iter`, engine_pretrain.py:163, into the seed).
        Follow with finish() as usual."""
        self._prescaled = self.world > 1
        seed = float(scale) / self.world
        loss.backward(gradient=torch.full_like(loss, seed) if seed != 1.0 else None)

    def finish(self):
        """Call after loss.backward(): joins the communication stream; on the discovery step performs the (non-overlapped)
        reduction and builds the buckets for the following steps."""
        discovery = self._discovering
        if self._discovering:
            self._build(self._order)
            self._discovering = False
            for b in self.buckets:
                for p, v in zip(b.params, b.views):
                    v.copy_(p.grad)
                    p.grad = v
                self._launch(b)
        if self.stream is not None:
            if self._stream_used:  # (a wait on a stream that took no part in a CUDA-graph capture would invalidate it)
                torch.cuda.current_stream(self.device).wait_stream(self.stream)
                self._stream_used = False
        else:
            for b in self.buckets:
                if b.work is not None:
                    b.work.wait()
                    if not self._prescaled:
                        b.flat.mul_(1.0 / self.world)
                    b.work = None
        if not discovery:
            if self.world > 1 and any(b.pending < 0 for b in self.buckets):
                # gradients kept arriving after a bucket had been reduced: a second backward() into live gradients without
                # no_sync() around the first — the sums in the buckets are part reduced, part local
                stuck = [n for b in self.buckets if b.pending < 0 for n in b.names][:4]
                for b in self.buckets:
                    b.pending, b.launched = len(b.params), False
                raise RuntimeError("GradReducer.finish(): some buckets were reduced before all of their gradients had arrived "
                                   "(wrap all but the last backward pass of an accumulation group in reducer.no_sync()); "
                                   f"first parameters: {stuck}")
            late = [b for b in self.buckets if not b.launched]
            # buckets whose hooks did not all fire: after no_sync() micro-steps the in-place weight-gradient sinks accumulate
            # without going through autograd (ops._sink), so the last pass cannot count them — reduce those buckets now
            for b in late:
                self._launch(b)
            if late and self.stream is not None and self._stream_used:
                torch.cuda.current_stream(self.device).wait_stream(self.stream)
                self._stream_used = False
            for b in late:
                if b.work is not None:
                    b.work.wait()
                    if not self._prescaled:
                        b.flat.mul_(1.0 / self.world)
                    b.work = None
        for b in self.buckets:
            b.pending, b.launched = len(b.params), False
        self._prescaled = False

    def zero_grad(self):
        """Keeps param.grad pointing into the buckets (so the next hooks need no copy when autograd accumulates in place)
        and zeroes them; equivalent to optimizer.zero_grad(set_to_none=False)."""
        if self.device.type == "cuda":
            from . import ops
            ops._sinks_pass_done()  # (a backward pass that died half-way never ran its end-of-pass callback)
        if self.buckets is None:
            self.model.zero_grad(set_to_none=True)
            return
        for b in self.buckets:
            for p in b.params:
                p.grad = None

    def reset(self):
        self._clear_sinks()
        self.buckets, self._slot, self._order, self._discovering = None, {}, [], True
        self._seen = set()
        self._sym = None

    def remove(self):
        self._clear_sinks()
        for h in self._hooks:
            h.remove()
        self._hooks = []
