"""OCTCube-IR contrastive step (BASELINE configs[4]) on N GPUs: two ViT-L towers -> 512-d projections -> L2 normalise ->
ClipLoss with the feature exchange folded into the loss kernels (octcubem_b200/clip.py) -> backward -> gradient exchange
(GradReducer).  OCT tower: the unmasked 3D ViT-L of open_clip/models_vit_st_flash_attn_nodrop.py (54 frames x 256 x 256, t_patch 3:
18 x 256 + 1 = 4609 tokens, head_dim 64); en-face / IR tower: ViT-L/16 at 224 px, built from the same class with one temporal
slot of 3 frames (= the Conv2d over 3 channels of the reference's 2D tower: same GEMM, same 197 tokens).  The reference recipe
(train_IR_512-MAE3D-nodrop-vit-large.sh) runs batch 32 per GPU with gradient checkpointing; without checkpointing the 4609-token
tower holds ~4.5 GB of activations per sample, so the per-GPU batch here is 8 (stated in the output).  Prints one JSON line.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/bench_clip.py [--batch 8]
"""
import argparse
import json
import os
import sys
from functools import partial

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from octcubem_b200 import clip, models_vit_st_flash_attn as V  # noqa: E402
from octcubem_b200.dp import GradReducer  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--frames", type=int, default=54)
ap.add_argument("--steps", type=int, default=10)
a = ap.parse_args()
world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(0)
kw = dict(patch_size=16, in_chans=1, num_classes=512, embed_dim=1024, depth=24, num_heads=16, mlp_ratio=4.0,
          norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), dropout=0.0, sep_pos_embed=True, cls_embed=True, global_pool=True,
          use_flash_attn=True, precision="bf16")
model = clip.CustomTextCLIP(V.VisionTransformer(num_frames=a.frames, t_patch_size=3, img_size=256, **kw),
                            V.VisionTransformer(num_frames=3, t_patch_size=3, img_size=224, **kw)).to(dev)
if world > 1:
    for p in model.parameters():
        dist.broadcast(p.data, 0)
crit = clip.ClipLoss(local_loss=True, gather_with_grad=True, rank=rank, world_size=world)
reducer = GradReducer(model) if world > 1 else None
g = torch.Generator().manual_seed(100 + rank)
oct_vol = torch.rand(a.batch, 1, a.frames, 256, 256, generator=g).to(dev)
ir_img = torch.rand(a.batch, 1, 3, 224, 224, generator=g).to(dev)
loss_out = torch.zeros((), device=dev)


def step():
    if reducer is not None:
        reducer.zero_grad()
    else:
        model.zero_grad(set_to_none=True)
    img_f, txt_f, scale = model(oct_vol, ir_img)
    loss = crit(img_f, txt_f, scale)
    if reducer is not None:
        reducer.backward(loss)
        reducer.finish()
    else:
        loss.backward()
    loss_out.copy_(loss.detach())


for _ in range(3):
    step()
torch.cuda.synchronize()
graph = None
try:
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
except Exception as e:  # noqa: BLE001
    if rank == 0:
        print(f"[bench_clip] graph capture failed ({type(e).__name__}: {e}); eager", file=sys.stderr)
    graph = None
    torch.cuda.synchronize()
run = graph.replay if graph is not None else step
for _ in range(2):
    run()
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(a.steps):
    run()
e.record()
torch.cuda.synchronize()
ms_t = torch.tensor([s.elapsed_time(e) / a.steps], device=dev)
if world > 1:
    dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
ms = float(ms_t)
So, Si = (a.frames // 3) * 256 + 1, 197
gf = 3.0 * a.batch * (24 * (24 * So * 1024 ** 2 + 4 * So ** 2 * 1024) + 24 * (24 * Si * 1024 ** 2 + 4 * Si ** 2 * 1024)) / 1e9
if rank == 0:
    print(json.dumps({"workload": f"OCTCube-IR contrastive step (BASELINE configs[4]): per GPU {a.batch} x ({a.frames}x256x256 OCT volume, 3x224x224 "
                                  "en-face image), two ViT-L towers (S = %d / %d), 512-d L2-normalised features, ClipLoss with the feature "
                                  "exchange inside the loss kernels, backward, gradient exchange; bf16" % (So, Si),
                      "n_gpus": world, "pairs_per_s": world * a.batch / (ms / 1e3), "ms_per_step": ms, "cuda_graph": graph is not None,
                      "tower_tflops_per_gpu": gf / ms, "loss": float(loss_out), "peer_timeout": crit.peer_timeout(),
                      "allreduce": reducer.allreduce_backend() if reducer is not None else "none"}))
del graph
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
crit.close()
if reducer is not None:
    reducer.remove()
if world > 1:
    dist.destroy_process_group()
