"""Phase trace of the tcgen05 GEMM (build with tools/build_variant.sh gtrace "-DOCT_GEMM_TRACE=1" gemm_tc and run with
OCT_LIB=octcubem_b200/variants/libgtrace.so): per shape, CUDA-event time of a plain launch, then ONE launch with
OCT_GEMM_DBG=2 whose CTAs 0 / 1 / middle / last print their clock64 stamps (relative to the CTA's first instruction)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from octcubem_b200 import ops  # noqa: E402
from octcubem_b200._lib import EPI_BIAS, EPI_BIAS_GELU, EPI_DGELU, GEMM_NN, GEMM_NT, OCT_BF16  # noqa: E402

dev = torch.device("cuda:0")
shapes = [("enc proj fwd", GEMM_NT, 3280, 1024, 1024, EPI_BIAS), ("enc fc2 fwd", GEMM_NT, 3280, 1024, 4096, EPI_BIAS),
          ("enc qkv fwd", GEMM_NT, 3280, 3072, 1024, EPI_BIAS), ("enc fc1+gelu", GEMM_NT, 3280, 4096, 1024, EPI_BIAS_GELU),
          ("enc proj dgrad", GEMM_NN, 3280, 1024, 1024, 0),
          ("dec proj fwd", GEMM_NT, 32776, 512, 512, EPI_BIAS), ("dec fc1+gelu", GEMM_NT, 32776, 2048, 512, EPI_BIAS_GELU),
          ("dec fc2 fwd", GEMM_NT, 32776, 512, 2048, EPI_BIAS), ("dec fc2 dgrad+dgelu", GEMM_NN, 32776, 2048, 512, EPI_DGELU)]
only = sys.argv[1:]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for (name, lay, M, N, K, epi) in shapes:
    if only and not any(o in name for o in only):
        continue
    a = torch.randn(M, K, device=dev).bfloat16()
    b = torch.randn((N, K) if lay == GEMM_NT else (K, N), device=dev).bfloat16()
    bias = torch.randn(N, device=dev)
    out = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
    aux = torch.randn(M, N, device=dev).bfloat16() if epi in (EPI_BIAS_GELU, EPI_DGELU) else None
    fn = lambda: ops.gemm(lay, a, b, M, N, K, torch.bfloat16, epi, bias=bias if epi in (EPI_BIAS, EPI_BIAS_GELU) else None,
                          aux=aux, out=out, compute=OCT_BF16)
    os.environ["OCT_GEMM_DBG"] = "0"
    for _ in range(3):
        fn()
    ts = []
    for _ in range(10):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    ts.sort()
    # back-to-back launches (what the step's graph sees: no launch gap, warm L2)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(20):
        fn()
    e.record(); torch.cuda.synchronize()
    print(f"== {name}: M={M} N={N} K={K}: {ts[len(ts) // 2]:.1f} us flushed, {s.elapsed_time(e) * 1e3 / 20:.1f} us back to back", flush=True)
    flush.zero_()
    os.environ["OCT_GEMM_DBG"] = "2"
    fn()
    torch.cuda.synchronize()
    os.environ["OCT_GEMM_DBG"] = "0"
