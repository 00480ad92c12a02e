"""Times add+LN forward / backward at the step's two shapes (CUDA events, L2 flushed between launches)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from octcubem_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for (M, C) in ((32776, 512), (3280, 1024)):
    h = torch.randn(M, C, device=dev).bfloat16(); res = torch.randn(M, C, device=dev)
    gamma, beta = torch.randn(C, device=dev), torch.randn(C, device=dev)
    y, r, mean, rstd = ops.add_ln_fwd(h, res, gamma, beta, 1e-6, torch.bfloat16, True)
    dy = torch.randn(M, C, device=dev).bfloat16(); dres = torch.randn(M, C, device=dev)
    for name, fn, nbytes in (("fwd", lambda: ops.add_ln_fwd(h, res, gamma, beta, 1e-6, torch.bfloat16, True), 12),
                             ("bwd", lambda: ops.add_ln_bwd(dy, r, mean, rstd, gamma, dres, True, True), 16)):
        for _ in range(3):
            fn()
        ts = []
        for _ in range(10):
            flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); fn(); e.record(); torch.cuda.synchronize()
            ts.append(s.elapsed_time(e) * 1e3)
        ts.sort()
        us = ts[len(ts) // 2]
        print(f"add_ln {name} M={M} C={C}: {us:.1f} us  {M * C * nbytes / us / 1e3:.0f} GB/s algorithmic")
