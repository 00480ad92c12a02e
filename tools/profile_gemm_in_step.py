"""Per-launch GEMM durations inside a warm eager step, joined with the (layout, M, N, K, epilogue) recorded by a thin
wrapper around ops.gemm, grouped by shape."""
import collections
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from octcubem_b200 import models_mae, ops  # noqa: E402

calls = []
_orig = ops.gemm


def rec(layout, A, B, M, N, K, out_dtype, epilogue=0, **kw):
    if A.dtype == torch.bfloat16:
        calls.append((layout, M, N, K, epilogue, str(out_dtype)[6:]))
    return _orig(layout, A, B, M, N, K, out_dtype, epilogue, **kw)


ops.gemm = rec
dev = torch.device("cuda:0")
torch.manual_seed(0)
m = models_mae.flash_attn_mae_vit_large_patch16(input_size=256, in_chans=1, num_frames=48, t_patch_size=3, pred_t_dim=48,
                                                sep_pos_embed=True, cls_embed=True, high_res_input_size=512,
                                                decoder_embed_dim=512, decoder_depth=8, decoder_num_heads=16).to(dev)
vol = torch.rand(8, 1, 48, 256, 256, device=dev)


def step():
    m.zero_grad(set_to_none=True)
    loss, _, _ = m(vol, mask_ratio=0.9)
    loss.backward()


for _ in range(3):
    step()
torch.cuda.synchronize()
calls.clear()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
evs = sorted([e for e in prof.events() if "gemm_tc_kernel" in e.name], key=lambda e: e.time_range.start)
assert len(evs) == len(calls), (len(evs), len(calls))
agg = collections.defaultdict(lambda: [0, 0.0])
for e, c in zip(evs, calls):
    agg[c][0] += 1
    agg[c][1] += e.device_time
tot = sum(v[1] for v in agg.values())
print(f"GEMM total {tot / 1e3:.2f} ms")
names = {0: "NT", 1: "NN", 2: "TN"}
for c, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    layout, M, N, K, epi, od = c
    fl = 2.0 * M * N * K
    print(f"{t:8.1f} us  n={n:3d} avg={t / n:7.1f} us  {fl * n / t / 1e6:7.0f} TF/s  {names[layout]} M={M} N={N} K={K} epi={epi} {od}")
