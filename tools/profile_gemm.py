"""Runs the tcgen05 GEMM at step shapes (for ncu)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from octcubem_b200 import ops  # noqa: E402
from octcubem_b200._lib import EPI_BIAS, EPI_BIAS_GELU, EPI_DGELU, GEMM_NN, GEMM_NT, GEMM_TN, OCT_BF16  # noqa: E402

dev = torch.device("cuda:0")
shapes = [(GEMM_NT, 32776, 1536, 512, EPI_BIAS), (GEMM_NT, 32776, 2048, 512, EPI_BIAS_GELU), (GEMM_NT, 3280, 3072, 1024, EPI_BIAS),
          (GEMM_NN, 32776, 512, 2048, 0), (GEMM_TN, 2048, 512, 32776, 0), (GEMM_NN, 32776, 2048, 512, EPI_DGELU)]
if len(sys.argv) > 1:  # e.g. "1,5": only these entries
    shapes = [shapes[int(i)] for i in sys.argv[1].split(",")]
for (layout, M, N, K, epi) in shapes:
    a = torch.randn((K, M) if layout == GEMM_TN else (M, K), device=dev).bfloat16()
    b = torch.randn((N, K) if layout == GEMM_NT else (K, N), device=dev).bfloat16()
    bias = torch.randn(N, device=dev)
    od = torch.float32 if layout == GEMM_TN else torch.bfloat16
    aux = (torch.randn(M, N, device=dev).to(od) if epi == EPI_DGELU else torch.empty(M, N, dtype=od, device=dev)) if epi in (EPI_BIAS_GELU, EPI_DGELU) else None
    for _ in range(3):
        ops.gemm(layout, a, b, M, N, K, od, epi, bias=bias if epi in (EPI_BIAS, EPI_BIAS_GELU) else None, aux=aux, compute=OCT_BF16)
torch.cuda.synchronize()
print("done")
