"""CPU, world_size 2, gloo: the data-parallel gradient reducer (octcubem_b200/dp.py) reproduces DDP semantics —
N-rank averaged gradients == 1-rank gradients on the concatenated batch (SURVEY §4 'distributed', quirk Q12) —
skips parameters that take no part in the step (quirk Q13) and orders buckets by backward readiness."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from octcubem_b200.dp import GradReducer


class Net(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.front = torch.nn.Linear(16, 64)
        self.unused = torch.nn.Linear(16, 64)       # like high_res_patch_embed on a 3D-only step
        self.blocks = torch.nn.ModuleList([torch.nn.Linear(64, 64) for _ in range(6)])
        self.head = torch.nn.Linear(64, 8)

    def forward(self, x):
        x = torch.tanh(self.front(x))
        for b in self.blocks:
            x = x + torch.tanh(b(x))
        return self.head(x)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    net = Net()
    red = GradReducer(net, bucket_mb=0.02, first_bucket_mb=0.005, last_bucket_mb=0.006)
    g = torch.Generator().manual_seed(1)
    X, Y = torch.randn(3, 8, 16, generator=g), torch.randn(3, 8, 8, generator=g)   # 3 steps, global batch 8
    out = []
    for step in range(3):
        red.zero_grad()
        x, y = X[step].chunk(world)[rank], Y[step].chunk(world)[rank]
        loss = ((net(x) - y) ** 2).mean()          # LOCAL mean (Q12)
        if step == 1:
            red.backward(loss)                     # seeded with 1/world: no scaling pass after the collective
        else:
            loss.backward()
        red.finish()
        out.append({k: p.grad.clone() for k, p in net.named_parameters() if p.grad is not None})
    if rank == 0:
        torch.save((out, red.bucket_layout()), path)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_reducer_world2_matches_single_process(tmp_path):
    port = _free_port()
    ctx = mp.get_context("spawn")
    path = str(tmp_path / "rank0.pt")
    procs = [ctx.Process(target=_worker, args=(r, 2, port, path)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(90)
        assert p.exitcode == 0
    got, layout = torch.load(path)
    # single-process reference on the concatenated batch
    torch.manual_seed(0)
    net = Net()
    g = torch.Generator().manual_seed(1)
    X, Y = torch.randn(3, 8, 16, generator=g), torch.randn(3, 8, 8, generator=g)
    for step in range(3):
        net.zero_grad(set_to_none=True)
        ((net(X[step]) - Y[step]) ** 2).mean().backward()
        ref = {k: p.grad for k, p in net.named_parameters() if p.grad is not None}
        assert set(ref) == set(got[step]) and not any(k.startswith("unused") for k in got[step])
        for k in ref:
            assert torch.allclose(got[step][k], ref[k], rtol=1e-5, atol=1e-6), (step, k)
    # bucket layout: backward order (head first, front last), several buckets, small tail holding the last-ready params
    names = [n for b, _ in layout for n in b]
    assert names[0].startswith("head") and names[-1].startswith("front")
    assert len(layout) >= 3
    assert layout[-1][1] * 4 <= 0.006 * 2 ** 20 or len(layout[-1][0]) == 1
    assert sum(n for _, n in layout) == sum(p.numel() for k, p in Net().named_parameters() if not k.startswith("unused"))


def test_reducer_single_process_is_identity():
    torch.manual_seed(0)
    net = Net()
    red = GradReducer(net)
    x, y = torch.randn(4, 16), torch.randn(4, 8)
    for _ in range(2):
        red.zero_grad()
        ((net(x) - y) ** 2).mean().backward()
        red.finish()
    ref = Net()
    ref.load_state_dict(net.state_dict())
    ((ref(x) - y) ** 2).mean().backward()
    for (k, p), (_, r) in zip(net.named_parameters(), ref.named_parameters()):
        if r.grad is not None:
            assert torch.allclose(p.grad, r.grad, rtol=1e-6, atol=1e-7), k


def _worker_accumulate(rank, world, port, path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    net = Net()
    red = GradReducer(net, bucket_mb=0.02, first_bucket_mb=0.005, last_bucket_mb=0.006)
    x, y = torch.randn(4, 16), torch.randn(4, 8)
    red.zero_grad()
    ((net(x) - y) ** 2).mean().backward()
    red.finish()                                   # discovery step
    red.zero_grad()
    ((net(x) - y) ** 2).mean().backward()
    ((net(x) - y) ** 2).mean().backward()          # second pass into live gradients: buckets were already reduced
    msg = ""
    try:
        red.finish()
    except RuntimeError as e:
        msg = str(e)
    red.zero_grad()                                # the reducer stays usable afterwards
    ((net(x) - y) ** 2).mean().backward()
    red.finish()
    if rank == 0:
        torch.save(msg, path)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_reducer_refuses_gradient_accumulation_across_ranks(tmp_path):
    """Two backward passes before finish() at world size 2: the second pass's gradients were never reduced — finish() must
    say so instead of handing back half-reduced gradients (dp.GradReducer.finish)."""
    port = _free_port()
    ctx = mp.get_context("spawn")
    path = str(tmp_path / "msg.pt")
    procs = [ctx.Process(target=_worker_accumulate, args=(r, 2, port, path)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(90)
        assert p.exitcode == 0
    assert "not reduced" in torch.load(path)
