"""Contrastive step on the GPU against oracle/clip_loss_oracle.py (pinned to the unmodified reference ClipLoss on
2 gloo ranks by tests/test_clip_oracle.py): L2 normalisation, the fused exchange + logits + cross-entropy kernels at world size
1, and the multi-rank protocol with W emulated ranks on ONE GPU — every rank gets its own exchange buffer, state and stream, the
kernels of different ranks run concurrently and wait for each other's epoch flags exactly as they do over NVLink (the 2-/8-GPU
run of the same code is tools/check_clip_gpu.py)."""
import ctypes

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from octcubem_b200 import _lib, clip  # noqa: E402
from octcubem_b200.ops import _p  # noqa: E402
from oracle import clip_loss_oracle as OC  # noqa: E402

DEV = "cuda:0"


def rel(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B,D", [(32, 512), (5, 36), (3, 1024)])
def test_l2_normalize(dtype, B, D):
    g = torch.Generator().manual_seed(B)
    x = (torch.randn(B, D, generator=g) * 3).to(dtype)
    x[0] = 0                                                        # the eps-clamped row: y = 0, dx = dy / eps
    dy = torch.randn(B, D, generator=g)
    xr = x.float().clone().requires_grad_(True)
    F.normalize(xr, dim=-1).backward(dy)
    xd = x.to(DEV).requires_grad_(True)
    y = clip.l2_normalize(xd)
    assert y.dtype == torch.float32 and rel(y, F.normalize(x.float(), dim=-1)) < 1e-6
    y.backward(dy.to(DEV))
    assert rel(xd.grad[1:].float(), xr.grad[1:]) < (1e-5 if dtype == torch.float32 else 4e-3)
    assert torch.isfinite(xd.grad).all()


def _features(W, B, D, seed):
    g = torch.Generator().manual_seed(seed)
    f = [F.normalize(torch.randn(B, D, generator=g), dim=-1) for _ in range(2 * W)]
    return f[:W], f[W:]


@pytest.mark.parametrize("B,D", [(32, 512), (6, 32), (50, 128)])
def test_clip_loss_world1_vs_oracle(B, D):
    images, enfaces = _features(1, B, D, 1)
    scale = torch.tensor(14.285714)
    (loss_ref, d_img, d_enf, d_scale), = OC.clip_loss_and_grads(images, enfaces, scale)
    crit = clip.ClipLoss()
    img = images[0].to(DEV).requires_grad_(True)
    enf = enfaces[0].to(DEV).requires_grad_(True)
    sc = scale.to(DEV).requires_grad_(True)
    for _ in range(3):                                              # three epochs: both parities of the exchange buffer
        img.grad = enf.grad = sc.grad = None
        loss = crit(img, enf, sc)
        (loss * 1.5).backward()
        assert abs(float(loss) - float(loss_ref)) < 2e-6 * abs(float(loss_ref))
        assert rel(img.grad, 1.5 * d_img) < 2e-5 and rel(enf.grad, 1.5 * d_enf) < 2e-5
        assert abs(float(sc.grad) - 1.5 * float(d_scale)) < 2e-5 * abs(1.5 * float(d_scale)) + 1e-9
    assert not crit.peer_timeout()
    crit.close()


@pytest.mark.parametrize("W,B,D", [(2, 32, 512), (4, 6, 64), (6, 32, 512)])
def test_clip_loss_multi_rank_protocol_on_one_gpu(W, B, D):
    """W ranks emulated on one device, one stream each, launched in an order that forces every kernel to wait for flags raised
    by kernels launched AFTER it on other streams; three steps with fresh features (epoch parity, flag reuse)."""
    lib = _lib.load()
    scale = torch.tensor(14.285714)
    xb = [torch.zeros(lib.oct_clip_xchg_bytes(B, D) // 4, dtype=torch.int32, device=DEV) for _ in range(W)]
    table = (ctypes.c_void_p * W)(*[t.data_ptr() for t in xb])
    state = [torch.zeros(lib.oct_clip_state_bytes(B) // 4, dtype=torch.int32, device=DEV) for _ in range(W)]
    streams = [torch.cuda.Stream() for _ in range(W)]
    sc = scale.to(DEV)
    one = torch.ones((), device=DEV)
    for step in range(3):
        images, enfaces = _features(W, B, D, 10 + step)
        ref = OC.clip_loss_and_grads(images, enfaces, scale)
        img = [x.to(DEV) for x in images]
        enf = [x.to(DEV) for x in enfaces]
        loss = [torch.empty((), device=DEV) for _ in range(W)]
        d_img = [torch.empty(B, D, device=DEV) for _ in range(W)]
        d_enf = [torch.empty(B, D, device=DEV) for _ in range(W)]
        d_sc = [torch.empty((), device=DEV) for _ in range(W)]
        torch.cuda.synchronize()
        for r in range(W):
            with torch.cuda.stream(streams[r]):
                st = ctypes.c_void_p(streams[r].cuda_stream)
                rc = lib.oct_clip_loss_fwd(_p(img[r]), _p(enf[r]), _p(sc), table, _p(state[r]), _p(loss[r]), r, W, B, D, st)
                assert rc == 0, _lib.last_error()
                rc = lib.oct_clip_loss_bwd(_p(img[r]), _p(enf[r]), _p(sc), _p(one), table, _p(state[r]), _p(d_img[r]), _p(d_enf[r]),
                                           _p(d_sc[r]), r, W, B, D, st)
                assert rc == 0, _lib.last_error()
        torch.cuda.synchronize()
        for r in range(W):
            assert int(state[r][3]) == 0, "a kernel timed out waiting for a peer"
            assert int(state[r][0]) == step + 1
            l_ref, gi, ge, gs = ref[r]
            assert abs(float(loss[r]) - float(l_ref)) < 3e-6 * abs(float(l_ref)), (step, r)
            assert rel(d_img[r], gi) < 3e-5 and rel(d_enf[r], ge) < 3e-5, (step, r)
            assert abs(float(d_sc[r]) - float(gs)) < 3e-5 * abs(float(gs)) + 1e-9, (step, r)


def test_clip_loss_graph_replay():
    """The epoch lives on the device: one captured forward + backward replays as consecutive steps."""
    B, D = 32, 512
    crit = clip.ClipLoss()
    img = torch.zeros(B, D, device=DEV, requires_grad=True)
    enf = torch.zeros(B, D, device=DEV, requires_grad=True)
    sc = torch.tensor(14.285714, device=DEV, requires_grad=True)
    gi, ge, gs, lo = torch.zeros(B, D, device=DEV), torch.zeros(B, D, device=DEV), torch.zeros((), device=DEV), torch.zeros((), device=DEV)

    def step():
        loss = crit(img, enf, sc)
        a, b, c = torch.autograd.grad(loss, (img, enf, sc))
        gi.copy_(a); ge.copy_(b); gs.copy_(c); lo.copy_(loss.detach())

    images, enfaces = _features(1, B, D, 3)
    with torch.no_grad():
        img.copy_(images[0]); enf.copy_(enfaces[0])
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        step()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        step()
    for k in range(4):
        images, enfaces = _features(1, B, D, 20 + k)
        with torch.no_grad():
            img.copy_(images[0]); enf.copy_(enfaces[0])
        graph.replay()
        (l_ref, d_img, d_enf, d_scale), = OC.clip_loss_and_grads(images, enfaces, torch.tensor(14.285714))
        assert abs(float(lo) - float(l_ref)) < 2e-6 * abs(float(l_ref)), k
        assert rel(gi, d_img) < 2e-5 and rel(ge, d_enf) < 2e-5
        assert abs(float(gs) - float(d_scale)) < 2e-5 * abs(float(d_scale)) + 1e-9
    crit.close()


def test_clip_argument_validation():
    lib = _lib.load()
    t = (ctypes.c_void_p * 1)(16)
    rc = lib.oct_clip_loss_fwd(ctypes.c_void_p(16), ctypes.c_void_p(16), ctypes.c_void_p(16), t, ctypes.c_void_p(16), ctypes.c_void_p(16),
                               0, 1, 8, 6, None)
    assert rc == -1 and "D % 4" in _lib.last_error()
    with pytest.raises(NotImplementedError):
        clip.ClipLoss(local_loss=False, gather_with_grad=False, world_size=2)
