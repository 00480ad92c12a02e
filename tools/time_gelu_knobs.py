"""dec fc1-shaped GEMM (32776 x 2048 x 512) with each epilogue, with and without its global stores (OCT_GEMM_DBG=1: timing only)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from octcubem_b200 import ops
from octcubem_b200._lib import EPI_BIAS, EPI_BIAS_GELU, EPI_DGELU, EPI_NONE, GEMM_NN, GEMM_NT, OCT_BF16
dev = torch.device("cuda:0")
M, N, K = 32776, 2048, 512
a = torch.randn(M, K, device=dev).bfloat16(); b = torch.randn(N, K, device=dev).bfloat16(); bias = torch.randn(N, device=dev)
out = torch.empty(M, N, dtype=torch.bfloat16, device=dev); aux = torch.randn(M, N, device=dev).bfloat16()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for dbg in ("0", "1"):
    os.environ["OCT_GEMM_DBG"] = dbg
    for name, epi in (("none", EPI_NONE), ("bias", EPI_BIAS), ("bias+gelu", EPI_BIAS_GELU), ("dgelu", EPI_DGELU)):
        fn = lambda: ops.gemm(GEMM_NT, a, b, M, N, K, torch.bfloat16, epi, bias=bias if epi in (EPI_BIAS, EPI_BIAS_GELU) else None,
                              aux=aux if epi in (EPI_BIAS_GELU, EPI_DGELU) else None, out=out, compute=OCT_BF16)
        for _ in range(3): fn()
        ts = []
        for _ in range(10):
            flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e) * 1e3)
        ts.sort()
        print(f"stores {'off' if dbg == '1' else 'on '}  {name:10s} {ts[5]:7.1f} us")
