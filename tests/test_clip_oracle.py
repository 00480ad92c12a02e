"""CPU, world_size 2, gloo: the contrastive-loss oracle (oracle/clip_loss_oracle.py) against the UNMODIFIED
reference ClipLoss (retinal-COEM/src/open_clip/loss.py) in the recipe's configuration (--local-loss --gather-with-grad), and
its closed-form gradients against autograd.  The reference half needs /root/reference (build container only)."""
import importlib.util
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import clip_loss_oracle as OC

REF_LOSS = os.path.join(os.environ.get("OCT_REFERENCE_ROOT", "/root/reference"), "retinal-COEM", "src", "open_clip", "loss.py")
B, D, W = 6, 32, 2


def _features(seed=0):
    g = torch.Generator().manual_seed(seed)
    f = [torch.nn.functional.normalize(torch.randn(B, D, generator=g), dim=-1) for _ in range(2 * W)]
    return f[:W], f[W:], torch.tensor(14.285714)          # logit_scale = exp(log(1/0.07))


def test_closed_form_gradients_match_autograd_single_process():
    """Emulates the 2-rank step in one process: summing every rank's loss and differentiating w.r.t. the leaf features is what
    the autograd-aware all_gather (backward = SUM reduce-scatter) computes rank by rank."""
    images, enfaces, scale = _features()
    images = [x.clone().requires_grad_(True) for x in images]
    enfaces = [x.clone().requires_grad_(True) for x in enfaces]
    scales = [scale.clone().requires_grad_(True) for _ in range(W)]       # each rank owns a replica of the parameter
    all_image, all_enface = torch.cat(images), torch.cat(enfaces)
    losses = [OC.clip_loss_local(images[r], enfaces[r], all_image, all_enface, scales[r], r) for r in range(W)]
    sum(losses).backward()
    for r, (loss, d_img, d_enf, d_scale) in enumerate(OC.clip_loss_and_grads([x.detach() for x in images],
                                                                           [x.detach() for x in enfaces], scale)):
        assert torch.allclose(loss, losses[r].detach(), rtol=1e-6)
        assert torch.allclose(d_img, images[r].grad, rtol=1e-5, atol=1e-8)
        assert torch.allclose(d_enf, enfaces[r].grad, rtol=1e-5, atol=1e-8)
        assert torch.allclose(d_scale, scales[r].grad, rtol=1e-5, atol=1e-8)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    spec = importlib.util.spec_from_file_location("octcube_ref_open_clip_loss", REF_LOSS)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    images, enfaces, scale = _features()
    img = images[rank].clone().requires_grad_(True)
    enf = enfaces[rank].clone().requires_grad_(True)
    sc = scale.clone().requires_grad_(True)
    loss = ref.ClipLoss(local_loss=True, gather_with_grad=True, rank=rank, world_size=world)(img, enf, sc)
    loss.backward()
    torch.save((loss.detach(), img.grad, enf.grad, sc.grad), f"{path}.{rank}")
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
@pytest.mark.skipif(not os.path.isfile(REF_LOSS), reason="/root/reference not present")
def test_oracle_vs_reference_cliploss_two_ranks(tmp_path):
    port = _free_port()
    ctx = mp.get_context("spawn")
    path = str(tmp_path / "clip")
    procs = [ctx.Process(target=_worker, args=(r, W, port, path)) for r in range(W)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(90)
        assert p.exitcode == 0
    images, enfaces, scale = _features()
    want = OC.clip_loss_and_grads(images, enfaces, scale)
    for r in range(W):
        loss, d_img, d_enf, d_scale = torch.load(f"{path}.{r}")
        assert torch.allclose(loss, want[r][0], rtol=1e-6)
        assert torch.allclose(d_img, want[r][1], rtol=1e-5, atol=1e-8)
        assert torch.allclose(d_enf, want[r][2], rtol=1e-5, atol=1e-8)
        assert torch.allclose(d_scale, want[r][3], rtol=1e-5, atol=1e-8)
