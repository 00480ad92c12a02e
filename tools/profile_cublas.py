"""cuBLASLt (F.linear / torch.matmul) next to ours at a few step GEMM shapes, a few launches each — to be run under ncu:
kernel names (tile shape, cluster), grid, registers, shared memory, tensor-pipe / shared-memory utilisation of the incumbent."""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from octcubem_b200 import ops  # noqa: E402
from octcubem_b200._lib import EPI_BIAS, EPI_NONE, GEMM_NN, GEMM_NT, OCT_BF16  # noqa: E402

dev = torch.device("cuda:0")
for (M, N, K, kind) in ((32776, 512, 2048, "NT"), (3280, 1024, 1024, "NT"), (3280, 1024, 4096, "NT"), (32776, 512, 1536, "NN")):
    a = torch.randn(M, K, device=dev).bfloat16()
    w = torch.randn((N, K) if kind == "NT" else (K, N), device=dev).bfloat16()
    bias = torch.randn(N, device=dev)
    bias_lp = bias.bfloat16()
    out = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
    for _ in range(3):
        if kind == "NT":
            F.linear(a, w, bias_lp)
            ops.gemm(GEMM_NT, a, w, M, N, K, torch.bfloat16, EPI_BIAS, bias=bias, out=out, compute=OCT_BF16)
        else:
            torch.matmul(a, w)
            ops.gemm(GEMM_NN, a, w, M, N, K, torch.bfloat16, EPI_NONE, out=out, compute=OCT_BF16)
    torch.cuda.synchronize()
print("done")
