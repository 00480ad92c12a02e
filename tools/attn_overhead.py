"""Per-CTA fixed cost of the attention kernels: T(S) = CTAs/148 x (o + n_tiles x t) fitted over several sequence lengths."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from octcubem_b200 import ops
from octcubem_b200._lib import OCT_BF16
dev = torch.device("cuda:0")
H, d = 16, 32
def timeit(fn, n=10):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
rows = []
for B, S in ((8, 1025), (8, 2049), (8, 4097), (4, 8193), (8, 4096), (8, 4224)):
    qkv = (torch.randn(B, S, 3 * H * d, device=dev) * 0.5).bfloat16()
    dout = torch.randn(B, S, H * d, device=dev).bfloat16()
    out, lse = ops.attn_fwd(qkv, H, d, OCT_BF16)
    tf = timeit(lambda: ops.attn_fwd(qkv, H, d, OCT_BF16))
    tb = timeit(lambda: ops.attn_bwd(qkv, out, dout, lse, H, d, OCT_BF16))
    nkv = (S + 127) // 128
    print(f"B={B} S={S}: fwd {tf:8.1f} us  bwd {tb:8.1f} us | bwd CTAs {nkv * H * B} x {(S + 63) // 64} sub-tiles; fwd CTAs {((S + 255) // 256) * H * B} x {nkv} kv tiles")
    rows.append((B, S, tf, tb))
