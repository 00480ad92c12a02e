"""Runs every memory-bound kernel of the step once at its in-step shape (BASELINE cfg-2: B=8, 48x256x256) between
cudaProfilerStart/Stop, for ncu:
   ncu --profile-from-start off --set full --clock-control none -o gpurun_out/membound_r1 python tools/profile_membound.py
   python tools/ncu_membound_summary.py gpurun_out/membound_r1.ncu-rep > profiles/r1_membound_ncu.md
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from octcubem_b200 import ops, optim  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
B, T, L, keep = 8, 48, 4096, 409
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def work():
    out = []
    for (M, C) in ((B * (L + 1), 512), (B * (keep + 1), 1024)):                       # decoder / encoder residual streams
        h = torch.randn(M, C, device=dev).bfloat16(); res = torch.randn(M, C, device=dev)
        gamma, beta = torch.randn(C, device=dev), torch.randn(C, device=dev)
        flush.zero_()
        y, r, mean, rstd = ops.add_ln_fwd(h, res, gamma, beta, 1e-6, torch.bfloat16, True)
        dy = torch.randn(M, C, device=dev).bfloat16(); dres = torch.randn(M, C, device=dev)
        flush.zero_()
        out.append(ops.add_ln_bwd(dy, r, mean, rstd, gamma, dres, True, True))
    noise = torch.rand(B, L, device=dev)
    mask, ids_restore, ids_keep = ops.mask_sort(noise, keep)
    x = torch.randn(B, L, 1024, device=dev).bfloat16().requires_grad_(True)
    pos_sp = torch.randn(256, 1024, device=dev, requires_grad=True); pos_tmp = torch.randn(16, 1024, device=dev, requires_grad=True)
    cls = torch.randn(1024, device=dev, requires_grad=True)
    flush.zero_()
    g = ops.GatherTokensFn.apply(x, ids_keep, pos_sp, pos_tmp, cls)
    y = torch.randn(B, keep, 512, device=dev).bfloat16().requires_grad_(True)
    mt = torch.randn(512, device=dev, requires_grad=True)
    dsp = torch.randn(256, 512, device=dev, requires_grad=True); dtm = torch.randn(16, 512, device=dev, requires_grad=True)
    dcls = torch.randn(512, device=dev, requires_grad=True)
    flush.zero_()
    u = ops.UnshuffleFn.apply(y, ids_restore, mt, dsp, dtm, dcls)
    du = torch.randn_like(u)
    flush.zero_()
    u.backward(du)
    imgs = torch.rand(B, 1, T, 256, 256, device=dev)
    pred = torch.randn(B, L + 1, 768, device=dev).bfloat16().requires_grad_(True)
    flush.zero_()
    loss, _, _ = ops.MaskedMSELossFn.apply(imgs, pred, mask, 16, 3, 1, False, None)
    flush.zero_()
    loss.backward()
    cube = torch.randint(0, 256, (B, T, 256, 256), dtype=torch.uint8, device=dev)
    flush.zero_()
    out.append(ops.ingest_u8(cube, T))
    xs = torch.randn(2, 5121, 1024, device=dev).bfloat16()
    flush.zero_()
    out.append(ops.MeanPoolFn.apply(xs, 1, 5121, torch.float32))
    w = torch.randn(64 << 20, device=dev)
    flush.zero_()
    out.append(ops.cast_bf16(w))
    p = torch.nn.Parameter(w)
    p.grad = torch.randn_like(w)
    sh = torch.empty(w.shape, dtype=torch.bfloat16, device=dev)
    opt = optim.FusedAdamW([p], lr=1e-3, betas=(0.9, 0.95), shadows=lambda q: sh)
    flush.zero_()
    opt.step(max_grad_norm=1.0)
    return out


work()
torch.cuda.synchronize()
torch.cuda.profiler.start()
work()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
