"""Joint 2D-512 + 3D step (BASELINE cfg-4): what engine_pretrain.py:109-149 does per optimizer step —
forward_patch_embed (the engine's extra Conv3d pass), a 3D forward (frame_loss=True), a 2D-512 forward at its own mask
ratio, `loss = loss + loss_2d`, one backward — plus the fused AdamW update.  Prints one JSON line (informational; the
headline bench stays bench.py / cfg-2).
    python tools/bench_joint.py [--b3d 2] [--b2d 16] [--frames 60] [--steps 10]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from octcubem_b200 import models_mae, optim  # noqa: E402
from octcubem_b200.dp import GradReducer  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--b3d", type=int, default=2)
ap.add_argument("--b2d", type=int, default=16)
ap.add_argument("--frames", type=int, default=60)
ap.add_argument("--mask2d", type=float, default=0.75)
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--no-graph", action="store_true")
a = ap.parse_args()
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = models_mae.flash_attn_mae_vit_large_patch16(
    input_size=256, in_chans=1, num_frames=a.frames, t_patch_size=3, pred_t_dim=a.frames, sep_pos_embed=True, cls_embed=True,
    high_res_input_size=512, decoder_embed_dim=512, decoder_depth=8, decoder_num_heads=16, precision="bf16").to(dev)
vol = torch.rand(a.b3d, 1, a.frames, 256, 256, device=dev)
img = torch.rand(a.b2d, 1, 3, 512, 512, device=dev)
opt = optim.FusedAdamW(optim.add_weight_decay(model, 0.05), lr=1e-6, betas=(0.9, 0.95), shadows=model.shadow_of)
# world size 1: the reducer only pins every gradient to a fixed address (its flat buckets), which is what lets the whole
# step — optimizer table included — be captured into one CUDA graph
reducer = GradReducer(model)
total_out = torch.zeros((), device=dev)


def step():
    reducer.zero_grad()
    feat = model.forward_patch_embed(vol).detach()             # engine_pretrain.py:112 (feeds the dead get_mask pass)
    (loss, frame_loss), _, _ = model(vol, mask_ratio=0.9, frame_loss=True)
    loss_2d, _, _ = model(img, mask_ratio=a.mask2d)
    total = loss + loss_2d
    total.backward()
    reducer.finish()
    opt.step(max_grad_norm=1.0)
    model.shadows_current()
    total_out.copy_(total.detach())
    return feat


for _ in range(3):
    step()
torch.cuda.synchronize()
unused = sorted(k for k, p in model.named_parameters() if p.grad is None)
graph = None
if not a.no_graph:
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        step()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        step()
    graph.replay()
    torch.cuda.synchronize()
run = graph.replay if graph is not None else step
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(a.steps):
    run()
e.record()
torch.cuda.synchronize()
ms = s.elapsed_time(e) / a.steps
total = total_out
print(json.dumps({"workload": f"joint step: {a.b3d} x {a.frames}x256x256 volumes @0.9 + {a.b2d} x 3x512x512 en-face triplets @{a.mask2d}, "
                              "bf16 fwd+bwd + grad-norm clip + fused AdamW, " + ("one CUDA graph" if graph is not None else "eager"),
                  "ms_per_step": ms, "volumes_per_s": a.b3d / (ms / 1e3), "images_2d_per_s": a.b2d / (ms / 1e3),
                  "loss": float(total), "finite": bool(torch.isfinite(total)), "params_without_grad": unused}))
