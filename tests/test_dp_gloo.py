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
    """Two backward passes before finish() at world size 2 without no_sync(): the buckets were reduced after the first pass and
    the second pass added local gradients on top — finish() must say so instead of handing back half-reduced gradients."""
    port = _free_port()
    ctx = mp.get_context("spawn")
    path = str(tmp_path / "msg.pt")
    procs = [ctx.Process(target=_worker_accumulate, args=(r, 2, port, path)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(90)
        assert p.exitcode == 0
    assert "no_sync" in torch.load(path)


def _worker_no_sync(rank, world, port, path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    net = Net()
    red = GradReducer(net, bucket_mb=0.02, first_bucket_mb=0.005, last_bucket_mb=0.006)
    g = torch.Generator().manual_seed(3)
    X, Y = torch.randn(16, 16, generator=g), torch.randn(16, 8, generator=g)
    red.zero_grad()
    ((net(X[:2]) - Y[:2]) ** 2).mean().backward()
    red.finish()                                              # discovery step (its result is not used)
    out = {}
    for style in ("seeded", "plain"):
        red.zero_grad()
        micro = X.chunk(2), Y.chunk(2)                        # accumulation group of 2 micro-steps of global batch 8
        for k in range(2):
            x, y = micro[0][k].chunk(world)[rank], micro[1][k].chunk(world)[rank]
            loss = ((net(x) - y) ** 2).mean() / 2             # loss /= accum_iter (engine_pretrain.py:163)
            if k == 0:
                with red.no_sync():
                    red.backward(loss) if style == "seeded" else loss.backward()
            else:
                red.backward(loss) if style == "seeded" else loss.backward()
                red.finish()
        out[style] = {n: p.grad.clone() for n, p in net.named_parameters() if p.grad is not None}
    if rank == 0:
        torch.save(out, path)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_reducer_no_sync_accumulates_then_reduces(tmp_path):
    """reducer.no_sync() (DDP.no_sync's role): two micro-steps on two ranks == one pass over the 16 samples with the mean loss."""
    port = _free_port()
    ctx = mp.get_context("spawn")
    path = str(tmp_path / "acc.pt")
    procs = [ctx.Process(target=_worker_no_sync, args=(r, 2, port, path)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(90)
        assert p.exitcode == 0
    got = torch.load(path)
    torch.manual_seed(0)
    net = Net()
    g = torch.Generator().manual_seed(3)
    X, Y = torch.randn(16, 16, generator=g), torch.randn(16, 8, generator=g)
    ((net(X) - Y) ** 2).mean().backward()
    for style in ("seeded", "plain"):
        assert set(got[style]) == {n for n, p in net.named_parameters() if p.grad is not None}
        for n, p in net.named_parameters():
            if p.grad is not None:
                assert torch.allclose(got[style][n], p.grad, rtol=1e-5, atol=1e-7), (style, n)


class BranchNet(Net):
    """`unused` takes part only when asked to: on the last micro-step its hooks do not fire, like the in-place weight-gradient
    sinks of the CUDA path (ops._sink) that accumulate without going through autograd."""

    def forward(self, x, extra=False):
        h = torch.tanh(self.front(x))
        if extra:
            h = h + torch.tanh(self.unused(x))
        for b in self.blocks:
            h = h + torch.tanh(b(h))
        return self.head(h)


def _worker_late_bucket(rank, world, port, path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    net = BranchNet()
    red = GradReducer(net, bucket_mb=0.02, first_bucket_mb=0.005, last_bucket_mb=0.006)
    g = torch.Generator().manual_seed(4)
    X, Y = torch.randn(16, 16, generator=g), torch.randn(16, 8, generator=g)
    red.zero_grad()
    ((net(X[:2], extra=True) - Y[:2]) ** 2).mean().backward()
    red.finish()                                              # discovery with every parameter taking part
    red.zero_grad()
    for k in range(2):
        x, y = X.chunk(2)[k].chunk(world)[rank], Y.chunk(2)[k].chunk(world)[rank]
        loss = ((net(x, extra=(k == 0)) - y) ** 2).mean() / 2
        if k == 0:
            with red.no_sync():
                red.backward(loss)
        else:
            red.backward(loss)                                # `unused.*` gets no gradient in this pass: its bucket is late
            red.finish()
    if rank == 0:
        torch.save({n: p.grad.clone() for n, p in net.named_parameters()}, path)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_reducer_reduces_buckets_whose_hooks_did_not_fire_on_the_last_micro_step(tmp_path):
    port = _free_port()
    ctx = mp.get_context("spawn")
    path = str(tmp_path / "late.pt")
    procs = [ctx.Process(target=_worker_late_bucket, args=(r, 2, port, path)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(90)
        assert p.exitcode == 0
    got = torch.load(path)
    torch.manual_seed(0)
    net = BranchNet()
    g = torch.Generator().manual_seed(4)
    X, Y = torch.randn(16, 16, generator=g), torch.randn(16, 8, generator=g)
    (((net(X[:8], extra=True) - Y[:8]) ** 2).mean() / 2 + ((net(X[8:]) - Y[8:]) ** 2).mean() / 2).backward()
    for n, p in net.named_parameters():
        assert torch.allclose(got[n], p.grad, rtol=1e-5, atol=1e-7), n
