"""Joint 2D-512 + 3D step (BASELINE cfg-4): what engine_pretrain.py:83-173 does per optimizer step — a 3D
forward (frame_loss=True), a 2D-512 forward at its own mask ratio, `loss = loss + loss_2d`, one backward, grad-norm clip,
AdamW on the per-iteration cosine schedule — through octcubem_b200.engine_pretrain.JointPretrainStep (two eager steps,
then ONE CUDA graph per input signature; learning rate / step count live in the optimizer's device clock).
Prints one JSON line (informational; the headline bench stays bench.py / cfg-2).
    python tools/bench_joint.py [--b3d 2] [--b2d 16] [--frames 60] [--steps 10] [--no-graph]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/bench_joint.py ...
(N ranks: BASELINE configs[3]; per-rank batches, gradients exchanged by the reducer inside the captured step; rank 0 prints.)
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from octcubem_b200 import models_mae, optim  # noqa: E402
from octcubem_b200.engine_pretrain import JointPretrainStep  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--b3d", type=int, default=2)
ap.add_argument("--b2d", type=int, default=16)
ap.add_argument("--frames", type=int, default=60)
ap.add_argument("--mask2d", type=float, default=0.75)
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--no-graph", action="store_true")
a = ap.parse_args()
import torch.distributed as dist
world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(0)
model = models_mae.flash_attn_mae_vit_large_patch16(
    input_size=256, in_chans=1, num_frames=a.frames, t_patch_size=3, pred_t_dim=a.frames, sep_pos_embed=True, cls_embed=True,
    high_res_input_size=512, decoder_embed_dim=512, decoder_depth=8, decoder_num_heads=16, precision="bf16").to(dev)
if world > 1:
    for p_ in model.parameters():
        dist.broadcast(p_.data, 0)
torch.manual_seed(100 + rank)
vol = torch.rand(a.b3d, 1, a.frames, 256, 256, device=dev)
img = torch.rand(a.b2d, 1, 3, 512, 512, device=dev)
sched = optim.CosineSchedule(lr=1.6e-3, min_lr=1e-6, warmup_epochs=5, epochs=100, epochs_per_step=1.0 / 2000)
opt = optim.FusedAdamW(optim.add_weight_decay(model, 0.05), betas=(0.9, 0.95), shadows=model.shadow_of, schedule=sched)
engine = JointPretrainStep(model, opt, mask_ratio=0.9, clip_grad=1.0, use_graph=not a.no_graph, warm_steps=3)

for _ in range(5):                      # 3 eager steps, capture + first replay, one more replay
    res = engine(vol, img, mask_ratio_2d=a.mask2d)
torch.cuda.synchronize()
unused = sorted(k for k, p in model.named_parameters() if p.grad is None)
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(a.steps):
    res = engine(vol, img, mask_ratio_2d=a.mask2d)
e.record()
torch.cuda.synchronize()
ms_t = torch.tensor([s.elapsed_time(e) / a.steps], device=dev)
if world > 1:
    dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
ms = float(ms_t)
vals = res.check_finite()
step, lr = opt.clock_state()
graphed = any(ent["graph"] is not None for ent in engine._entries.values())
if rank == 0:
  print(json.dumps({"n_gpus": world, "allreduce": engine.reducer.allreduce_backend(),
                  "workload": f"joint step (BASELINE configs[3]), per GPU: {a.b3d} x {a.frames}x256x256 volumes @0.9 + {a.b2d} x 3x512x512 en-face triplets @{a.mask2d}, "
                              "bf16 fwd+bwd + grad-norm clip + fused AdamW on the device-clock cosine schedule, "
                              + ("one CUDA graph, inputs copied into its static buffers every step" if graphed else "eager"),
                  "ms_per_step": ms, "volumes_per_s": world * a.b3d / (ms / 1e3), "images_2d_per_s": world * a.b2d / (ms / 1e3),
                  "loss_all": vals["loss_all"], "grad_norm": vals.get("grad_norm"), "optimizer_steps": step, "lr": lr,
                  "lr_expected": sched.lr_at_step(step), "params_without_grad": unused}))
engine.close()
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
