"""torchrun check of dp.GradReducer on GPUs (symmetric-memory all-reduce of csrc/allreduce.cu, or NCCL with OCT_ALLREDUCE=nccl): the reduced gradients of a toy 3D-MAE step on `world` ranks must equal
the mean of the per-rank gradients computed locally without the reducer (DDP semantics), with the wgrad GEMMs writing
straight into the all-reduce buckets and the backward seeded with 1/world.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/check_dp_gpu.py
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from octcubem_b200 import models_mae, ops  # noqa: E402
from octcubem_b200.dp import GradReducer  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
# everything below runs on ONE non-default stream: autograd pins every AccumulateGrad node to the stream of its first backward
# pass, and a node created on the legacy default stream cannot take part in a later CUDA-graph capture
_main_stream = torch.cuda.Stream()
torch.cuda.set_stream(_main_stream)
kw = dict(input_size=64, patch_size=16, in_chans=1, embed_dim=128, depth=2, num_heads=2, decoder_embed_dim=64, decoder_depth=1,
          decoder_num_heads=2, num_frames=12, t_patch_size=3, pred_t_dim=12, high_res_input_size=128, sep_pos_embed=True,
          cls_embed=True, use_flash_attn=True, precision="bf16", norm_layer=lambda d: torch.nn.LayerNorm(d, eps=1e-6))
torch.manual_seed(0)
model = models_mae.MaskedAutoencoderViT(**kw).to(dev)
for p in model.parameters():
    dist.broadcast(p.data, 0)
g = torch.Generator().manual_seed(7)
vols = torch.rand(world, 2, 1, 12, 64, 64, generator=g).to(dev)
noises = torch.rand(world, 2, 4 * 16, generator=g).to(dev)

# reference: every rank's gradient computed locally, averaged
ref = None
for r in range(world):
    model.zero_grad(set_to_none=True)
    loss, _, _ = model(vols[r], mask_ratio=0.75, noise=noises[r])
    loss.backward()
    grads = {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}
    ref = grads if ref is None else {k: ref[k] + grads[k] for k in ref}
ref = {k: v / world for k, v in ref.items()}

red = GradReducer(model, bucket_mb=0.25, first_bucket_mb=0.05, last_bucket_mb=0.05)
worst = 0.0
for step in range(3):  # step 0 = discovery (non-overlapped), then the hooked / in-place path
    red.zero_grad()
    loss, _, _ = model(vols[rank], mask_ratio=0.75, noise=noises[rank])
    red.backward(loss)
    red.finish()
    torch.cuda.synchronize()
    for k, p in model.named_parameters():
        if p.grad is None:
            continue
        err = float((p.grad - ref[k]).norm() / (ref[k].norm() + 1e-20))
        worst = max(worst, err)
        assert err < 2e-3, (step, k, err)  # same bf16 kernels on both sides; only the split-K / reduction order differs
in_place = sum(1 for k, p in model.named_parameters() if p.grad is not None and p.data_ptr() in ops.grad_sinks
               and p.grad.data_ptr() == ops.grad_sinks[p.data_ptr()][0].data_ptr())
# a second pass through a CUDA graph (device-resident epochs of the symmetric all-reduce must survive replay)
def one():
    red.zero_grad()
    loss, _, _ = model(vols[rank], mask_ratio=0.75, noise=noises[rank])
    red.backward(loss)
    red.finish()


one()
torch.cuda.synchronize()
graph = torch.cuda.CUDAGraph()
with torch.cuda.graph(graph, stream=_main_stream):
    one()
for _ in range(3):
    graph.replay()
torch.cuda.synchronize()
for k, p in model.named_parameters():
    if p.grad is not None:
        err = float((p.grad - ref[k]).norm() / (ref[k].norm() + 1e-20))
        worst = max(worst, err)
        assert err < 2e-3, ("graph", k, err)
assert not red.peer_timeout(), "an all-reduce kernel timed out waiting for a peer"
print(f"rank {rank}: OK, worst rel err {worst:.2e}, {in_place} gradients live in their buckets, {len(red.buckets)} buckets, "
      f"all-reduce backend: {red.allreduce_backend()}", flush=True)
del graph
dist.barrier()
dist.destroy_process_group()
