"""One eager bf16 training step (BASELINE cfg-2: B=8, 48x256x256) between cudaProfilerStart/Stop, for ncu:
   ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file X python tools/profile_step.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from octcubem_b200 import models_mae  # noqa: E402

B = int(os.environ.get("OCT_PROFILE_BATCH", "8"))
T = int(os.environ.get("OCT_PROFILE_FRAMES", "48"))
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
    return loss


for _ in range(2):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
loss = step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("loss", float(loss))
