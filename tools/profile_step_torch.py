"""Per-kernel device times of warm eager training steps via torch.profiler (CUPTI): unlike the ncu launch list these are
not cold-cache / serialised."""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from octcubem_b200 import models_mae  # noqa: E402

B, T = 8, 48
dev = torch.device("cuda:0")
torch.manual_seed(0)
m = models_mae.flash_attn_mae_vit_large_patch16(input_size=256, in_chans=1, num_frames=T, t_patch_size=3, pred_t_dim=T,
                                                sep_pos_embed=True, cls_embed=True, high_res_input_size=512,
                                                decoder_embed_dim=512, decoder_depth=8, decoder_num_heads=16).to(dev)
vol = torch.rand(B, 1, T, 256, 256, device=dev)


def step():
    m.zero_grad(set_to_none=True)
    loss, _, _ = m(vol, mask_ratio=0.9)
    loss.backward()


for _ in range(3):
    step()
torch.cuda.synchronize()
N = 3
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(N):
        step()
    torch.cuda.synchronize()
rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
tot = sum(e.device_time_total for e in rows)
print(f"total device time per step: {tot / N / 1e3:.2f} ms")
for e in rows[:30]:
    print(f"{e.device_time_total / N:10.1f} us {100 * e.device_time_total / tot:5.1f}%  n={e.count // N:4d}  avg={e.device_time_total / e.count:8.1f}  {e.key[:90]}")
