"""N-rank check of the contrastive step over NVLink (BASELINE configs[4]); launch with torchrun, one rank per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 tools/check_clip_gpu.py

Every rank runs octcubem_b200.clip.ClipLoss (exchange buffers peer-mapped through CUDA IPC, features / lse vectors read from the
peers inside the kernels) on its own features, B = 32 per rank, D = 512, for several steps eagerly and replayed from a CUDA
graph, and compares loss, d image, d enface and d logit_scale with oracle/clip_loss_oracle.py evaluated on the features of all
ranks (collected with an ordinary all_gather — checker only).  Rank 0 prints one JSON line.  Measurement / test tool."""
import json
import os
import sys

import torch
import torch.distributed as dist
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from octcubem_b200 import clip  # noqa: E402
from oracle import clip_loss_oracle as OC  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    B, D = 32, 512
    crit = clip.ClipLoss(local_loss=True, gather_with_grad=True, rank=rank, world_size=world)
    scale = torch.tensor(14.285714)
    img = torch.zeros(B, D, device=dev, requires_grad=True)
    enf = torch.zeros(B, D, device=dev, requires_grad=True)
    sc = scale.to(dev).requires_grad_(True)
    gi, ge = torch.zeros(B, D, device=dev), torch.zeros(B, D, device=dev)
    gs, lo = torch.zeros((), device=dev), torch.zeros((), device=dev)

    def step():
        loss = crit(img, enf, sc)
        a, b, c = torch.autograd.grad(loss, (img, enf, sc))
        gi.copy_(a); ge.copy_(b); gs.copy_(c); lo.copy_(loss.detach())

    def feed(k):
        g = torch.Generator().manual_seed(1000 * k + rank)
        with torch.no_grad():
            img.copy_(F.normalize(torch.randn(B, D, generator=g), dim=-1))
            enf.copy_(F.normalize(torch.randn(B, D, generator=g), dim=-1))

    def check(k, tag):
        all_i = [torch.empty(B, D, device=dev) for _ in range(world)]
        all_e = [torch.empty(B, D, device=dev) for _ in range(world)]
        dist.all_gather(all_i, img.detach())
        dist.all_gather(all_e, enf.detach())
        ref = OC.clip_loss_and_grads([t.cpu() for t in all_i], [t.cpu() for t in all_e], scale)[rank]
        rel = lambda a, b: float((a.double().cpu() - b.double()).norm() / b.double().norm())  # noqa: E731
        errs = {"loss": abs(float(lo) - float(ref[0])) / abs(float(ref[0])), "d_image": rel(gi, ref[1]), "d_enface": rel(ge, ref[2]),
                "d_scale": abs(float(gs) - float(ref[3])) / (abs(float(ref[3])) + 1e-12)}
        worst = max(errs.values())
        t = torch.tensor([worst], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        assert float(t) < 5e-5, (tag, k, rank, errs)
        return float(t)

    worst = 0.0
    for k in range(3):                                   # eager steps
        feed(k)
        step()
        worst = max(worst, check(k, "eager"))
    torch.cuda.synchronize()
    dist.barrier()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        feed(3)
        step()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        step()
    for k in range(4, 8):                                # replayed steps
        feed(k)
        graph.replay()
        worst = max(worst, check(k, "graph"))
    # latency of the fused exchange + loss (forward + backward), all ranks in lock-step
    torch.cuda.synchronize()
    dist.barrier()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(50):
        graph.replay()
    e.record()
    torch.cuda.synchronize()
    us = torch.tensor([s.elapsed_time(e) * 1e3 / 50], device=dev)
    dist.all_reduce(us, op=dist.ReduceOp.MAX)
    # the library baseline of the same step: autograd-aware all_gather (NCCL) + two matmuls + two cross-entropies in torch
    import torch.distributed.nn as dnn
    def lib_step():
        ai = torch.cat(dnn.all_gather(img), 0)
        ae = torch.cat(dnn.all_gather(enf), 0)
        labels = torch.arange(B, device=dev) + B * rank
        loss = (F.cross_entropy(sc * img @ ae.T, labels) + F.cross_entropy(sc * enf @ ai.T, labels)) / 2
        torch.autograd.grad(loss, (img, enf, sc))
    for _ in range(3):
        lib_step()
    torch.cuda.synchronize()
    dist.barrier()
    s.record()
    for _ in range(20):
        lib_step()
    e.record()
    torch.cuda.synchronize()
    lib_us = torch.tensor([s.elapsed_time(e) * 1e3 / 20], device=dev)
    dist.all_reduce(lib_us, op=dist.ReduceOp.MAX)
    timeout = crit.peer_timeout()
    if rank == 0:
        print(json.dumps({"check": "clip loss fused exchange over peer memory", "world": world, "B_per_rank": B, "D": D,
                          "worst_rel_err_vs_oracle": worst, "peer_timeout": timeout, "fwd_bwd_us_graph": float(us),
                          "torch_nccl_all_gather_eager_us": float(lib_us), "steps_checked": 7}), flush=True)
    del graph
    torch.cuda.synchronize()
    dist.barrier()
    crit.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
