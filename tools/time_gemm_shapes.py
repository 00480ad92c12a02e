"""Times every distinct GEMM of one training step (BASELINE cfg-2: encoder 3280 tokens x 1024, decoder 32776 tokens x 512)
standalone with CUDA events; prints us and TFLOP/s per shape and the per-step total (x launches)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from octcubem_b200 import ops  # noqa: E402
from octcubem_b200._lib import EPI_BIAS, EPI_BIAS_GELU, EPI_DGELU, GEMM_NN, GEMM_NT, OCT_BF16  # noqa: E402

dev = torch.device("cuda:0")
Te, Td, E, D = 3280, 32776, 1024, 512
NT, NN, WG = "NT", "NN", "WG"
shapes = []
for tag, T, C, reps in (("enc", Te, E, 24), ("dec", Td, D, 8)):
    shapes += [(f"{tag} qkv fwd", NT, T, 3 * C, C, EPI_BIAS, reps), (f"{tag} proj fwd", NT, T, C, C, EPI_BIAS, reps),
               (f"{tag} fc1 fwd+gelu", NT, T, 4 * C, C, EPI_BIAS_GELU, reps), (f"{tag} fc2 fwd", NT, T, C, 4 * C, EPI_BIAS, reps),
               (f"{tag} qkv dgrad", NN, T, C, 3 * C, 0, reps), (f"{tag} proj dgrad", NN, T, C, C, 0, reps),
               (f"{tag} fc1 dgrad", NN, T, C, 4 * C, 0, reps), (f"{tag} fc2 dgrad+dgelu", NN, T, 4 * C, C, EPI_DGELU, reps),
               (f"{tag} qkv wgrad", WG, 3 * C, C, T, 0, reps), (f"{tag} proj wgrad", WG, C, C, T, 0, reps),
               (f"{tag} fc1 wgrad", WG, 4 * C, C, T, 0, reps), (f"{tag} fc2 wgrad", WG, C, 4 * C, T, 0, reps)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
tot = 0.0
for (name, kind, M, N, K, epi, reps) in shapes:
    if kind == WG:
        dy = torch.randn(K, M, device=dev).bfloat16(); x = torch.randn(K, N, device=dev).bfloat16()
        fn = lambda: ops.wgrad_bias(dy, x)
    else:
        a = torch.randn(M, K, device=dev).bfloat16()
        b = torch.randn((N, K) if kind == NT else (K, N), device=dev).bfloat16()
        bias = torch.randn(N, device=dev)
        out = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
        aux = torch.randn(M, N, device=dev).bfloat16() if epi in (EPI_BIAS_GELU, EPI_DGELU) else None
        lay = GEMM_NT if kind == NT else GEMM_NN
        fn = lambda: ops.gemm(lay, a, b, M, N, K, torch.bfloat16, epi, bias=bias if epi in (EPI_BIAS, EPI_BIAS_GELU) else None,
                              aux=aux, out=out, compute=OCT_BF16)
    for _ in range(3):
        fn()
    ts = []
    for _ in range(10):
        flush.zero_()  # L2 flush between timed launches
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    ts.sort()
    us = ts[len(ts) // 2]
    tot += us * reps
    print(f"{name:22s} {kind} M={M:6d} N={N:5d} K={K:6d}  {us:7.1f} us  {2.0 * M * N * K / us / 1e6:7.1f} TF/s  x{reps} = {us * reps / 1e3:6.2f} ms")
print(f"total {tot / 1e3:.2f} ms per step")
