"""Runs the tcgen05 attention fwd+bwd at the decoder (or encoder) shape of BASELINE cfg-2 a few times (for ncu)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from octcubem_b200 import ops  # noqa: E402
from octcubem_b200._lib import OCT_BF16  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "dec"
B, S, H, d = (8, 4097, 16, 32) if which == "dec" else (8, 410, 16, 64)
dev = torch.device("cuda:0")
qkv = (torch.randn(B, S, 3 * H * d, device=dev) * 0.5).bfloat16()
dout = torch.randn(B, S, H * d, device=dev).bfloat16()
for _ in range(3):
    out, lse = ops.attn_fwd(qkv, H, d, OCT_BF16)
    dq = ops.attn_bwd(qkv, out, dout, lse, H, d, OCT_BF16)
torch.cuda.synchronize()
print("done", float(dq.float().abs().mean()))
if len(sys.argv) > 2 and sys.argv[2] == "time":
    def timeit(fn, n=20):
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n
    print(f"[TIME] attn fwd {which}: {timeit(lambda: ops.attn_fwd(qkv, H, d, OCT_BF16)) * 1e3:.1f} us")
    print(f"[TIME] attn bwd {which}: {timeit(lambda: ops.attn_bwd(qkv, out, dout, lse, H, d, OCT_BF16)) * 1e3:.1f} us")
