"""Forward/backward error of the tcgen05 attention against fp64 torch for a list of sequence lengths, per 128-row block."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from octcubem_b200 import ops  # noqa: E402
from octcubem_b200._lib import OCT_BF16  # noqa: E402

dev = "cuda:0"
d = int(os.environ.get("HD", "32"))
for S in [int(a) for a in sys.argv[1:]]:
    B, H = 1, 3
    g = torch.Generator().manual_seed(S)
    qkv = torch.randn(B, S, 3 * H * d, generator=g).bfloat16()
    dout = torch.randn(B, S, H * d, generator=g).bfloat16()
    x = qkv.double().requires_grad_(True)
    q, k, v = x.view(B, S, 3, H, d).unbind(2)
    s = torch.einsum("bthd,bshd->bhts", q, k) / math.sqrt(d)
    ref = torch.einsum("bhts,bshd->bthd", torch.softmax(s, -1), v).reshape(B, S, H * d)
    (ref * dout.double()).sum().backward()
    qd = qkv.to(dev).requires_grad_(True)
    out = ops.AttnFn.apply(qd, H, OCT_BF16)
    out.backward(dout.to(dev))
    o = out.detach().double().cpu()
    gq = qd.grad.double().cpu()
    rel = lambda a, b: float((a - b).norm() / (b.norm() + 1e-30))
    blocks = " ".join(f"{rel(o[:, r:r + 128], ref.detach()[:, r:r + 128]):.4f}" for r in range(0, S, 128))
    gb = " ".join(f"{rel(gq[:, r:r + 128], x.grad[:, r:r + 128]):.4f}" for r in range(0, S, 128))
    print(f"S={S} d={d} fwd rel {rel(o, ref.detach()):.5f} | per block: {blocks}")
    print(f"        bwd rel {rel(gq, x.grad):.5f} | per block: {gb}")
