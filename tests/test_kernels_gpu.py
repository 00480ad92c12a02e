"""GPU parity tests of the individual kernels (through the C ABI) against CPU/torch fp32-or-better references."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from octcubem_b200 import ops  # noqa: E402
from octcubem_b200._lib import EPI_BIAS, GEMM_NN, GEMM_NT, GEMM_TN, OCT_BF16, OCT_F32  # noqa: E402
from oracle import mae3d_oracle as O  # noqa: E402

DEV = "cuda:0"


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


# ---------------------------------------------------------------- masking: bit-exact (integer work)
@pytest.mark.parametrize("L", [1024, 4096, 5120])
@pytest.mark.parametrize("kind", ["natural", "tiefree", "quantized"])
def test_mask_sort_golden(golden_dir, L, kind):
    g = np.load(os.path.join(golden_dir, "masking_cases.npz"))
    noise = torch.from_numpy(g[f"{kind}_{L}_noise"]).to(DEV)
    keep = g[f"{kind}_{L}_ids_keep"].shape[1]
    mask, ids_restore, ids_keep = ops.mask_sort(noise, keep)
    assert ids_restore.dtype == torch.int64 and ids_keep.dtype == torch.int64 and mask.dtype == torch.float32
    assert np.array_equal(ids_restore.cpu().numpy(), g[f"{kind}_{L}_ids_restore"].astype(np.int64))
    assert np.array_equal(ids_keep.cpu().numpy(), g[f"{kind}_{L}_ids_keep"].astype(np.int64))
    assert np.array_equal(mask.cpu().numpy().astype(np.uint8), g[f"{kind}_{L}_mask"])


@pytest.mark.parametrize("B,L,keep", [(1, 1, 0), (1, 1, 1), (3, 7, 3), (8, 5120, 511), (2, 16384, 1638), (64, 1024, 256), (2, 333, 333)])
def test_mask_sort_vs_oracle_and_torch_cuda(B, L, keep):
    g = torch.Generator().manual_seed(L + keep)
    noise = torch.rand(B, L, generator=g)
    noise[:, L // 3:] = torch.floor(noise[:, L // 3:] * 50) / 50  # plenty of ties
    mask, ids_restore, ids_keep = ops.mask_sort(noise.to(DEV), keep)
    sh = torch.argsort(noise, dim=1, stable=True)
    rs = torch.argsort(sh, dim=1, stable=True)
    assert torch.equal(ids_restore.cpu(), rs)
    assert torch.equal(ids_keep.cpu(), sh[:, :keep])
    assert torch.equal(mask.cpu(), (rs >= keep).float())
    # the contract of SURVEY H1: equals what the reference computes on ITS device (torch.argsort on CUDA, radix = stable)
    sh_cuda = torch.argsort(noise.to(DEV), dim=1)
    assert torch.equal(torch.argsort(sh_cuda, dim=1).cpu(), ids_restore.cpu())


def test_mask_sort_special_values_and_identity():
    noise = torch.tensor([[0.0, -0.0, float("inf"), -float("inf"), 1.0, -1.0, 0.0, 1.0]])
    mask, ids_restore, ids_keep = ops.mask_sort(noise.to(DEV), 3)
    sh = torch.argsort(noise, dim=1, stable=True)
    assert torch.equal(ids_restore.cpu(), torch.argsort(sh, dim=1, stable=True))
    ar = torch.arange(100, dtype=torch.float32).expand(2, 100).contiguous()  # mask_ratio == 0 path (models...:350-352)
    mask, ids_restore, ids_keep = ops.mask_sort(ar.to(DEV), 100)
    assert torch.equal(ids_restore.cpu(), torch.arange(100).expand(2, 100)) and float(mask.sum()) == 0


def test_mask_sort_empty():
    m, r, k = ops.mask_sort(torch.empty(0, 16, device=DEV), 4)
    assert r.shape == (0, 16) and k.shape == (0, 4)


# ---------------------------------------------------------------- patchify / gather / unshuffle (pure copies: exact)
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_patchify_exact(dtype):
    imgs = torch.rand(2, 1, 6, 64, 48)
    want = O.patchify(imgs, 16, 3)
    got = ops.patchify(imgs.to(DEV), 16, 3, dtype)
    assert torch.equal(got.cpu(), want.to(dtype))
    ids = torch.stack([torch.randperm(want.shape[1])[:5] for _ in range(2)])
    gk = ops.patchify(imgs.to(DEV), 16, 3, dtype, ids_keep=ids.to(DEV))
    assert torch.equal(gk.cpu(), torch.gather(want, 1, ids[..., None].expand(-1, -1, 768)).to(dtype))
    fidx = torch.tensor([0, 2, 5])
    gf = ops.patchify(imgs.to(DEV), 16, 3, dtype, frame_idx=fidx.to(DEV))
    assert torch.equal(gf.cpu(), O.patchify(imgs[:, :, fidx], 16, 3).to(dtype))


@pytest.mark.parametrize("xdtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("Tp", [1, 4])
def test_gather_tokens_fwd_bwd(xdtype, Tp):
    B, G, C, keep = 3, 16, 64, 7
    L = Tp * G
    g = torch.Generator().manual_seed(0)
    x = torch.randn(B, L, C, generator=g).to(xdtype)
    ids = torch.stack([torch.randperm(L, generator=g)[:keep] for _ in range(B)])
    sp = torch.randn(G, C, generator=g, requires_grad=True)
    tmp = torch.randn(Tp, C, generator=g, requires_grad=True) if Tp > 1 else None
    cls = torch.randn(C, generator=g, requires_grad=True)
    xr = x.float().clone().requires_grad_(True)
    pos = sp.repeat(Tp, 1) + (torch.repeat_interleave(tmp, G, dim=0) if tmp is not None else 0)
    want = torch.cat([cls.expand(B, 1, C), torch.gather(xr, 1, ids[..., None].expand(-1, -1, C)) + pos[ids]], 1)
    dout = torch.randn(B, keep + 1, C, generator=g)
    want.backward(dout)
    xd = x.detach().to(DEV).requires_grad_(True)
    spd, clsd = sp.detach().to(DEV).requires_grad_(True), cls.detach().to(DEV).requires_grad_(True)
    tmpd = tmp.detach().to(DEV).requires_grad_(True) if tmp is not None else None
    got = ops.GatherTokensFn.apply(xd, ids.to(DEV), spd, tmpd, clsd)
    assert rel(got, want.detach()) < 1e-6
    got.backward(dout.to(DEV))
    assert rel(spd.grad, sp.grad) < 1e-6 and rel(clsd.grad, cls.grad) < 1e-6
    if tmp is not None:
        assert rel(tmpd.grad, tmp.grad) < 1e-6
    assert rel(xd.grad.float(), xr.grad.to(xdtype).float()) < 1e-6


@pytest.mark.parametrize("ydtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("Tp", [1, 4])
def test_unshuffle_fwd_bwd(ydtype, Tp):
    B, G, D, keep = 2, 16, 32, 6
    L = Tp * G
    g = torch.Generator().manual_seed(1)
    y = torch.randn(B, keep, D, generator=g).to(ydtype)
    ids_restore = torch.stack([torch.randperm(L, generator=g) for _ in range(B)])
    mt = torch.randn(D, generator=g, requires_grad=True)
    sp = torch.randn(G, D, generator=g, requires_grad=True)
    tmp = torch.randn(Tp, D, generator=g, requires_grad=True) if Tp > 1 else None
    cls = torch.randn(D, generator=g, requires_grad=True)
    yr = y.float().clone().requires_grad_(True)
    x_ = torch.cat([yr, mt.expand(B, L - keep, D)], 1)
    x_ = torch.gather(x_, 1, ids_restore[..., None].expand(-1, -1, D))
    pos = sp.repeat(Tp, 1) + (torch.repeat_interleave(tmp, G, dim=0) if tmp is not None else 0)
    want = torch.cat([cls.expand(B, 1, D), x_ + pos], 1)
    dout = torch.randn(B, L + 1, D, generator=g)
    want.backward(dout)
    yd = y.detach().to(DEV).requires_grad_(True)
    leaves = [t.detach().to(DEV).requires_grad_(True) if t is not None else None for t in (mt, sp, tmp, cls)]
    got = ops.UnshuffleFn.apply(yd, ids_restore.to(DEV), leaves[0], leaves[1], leaves[2], leaves[3])
    assert rel(got, want.detach()) < 1e-6
    got.backward(dout.to(DEV))
    for a, b in zip(leaves, (mt, sp, tmp, cls)):
        if a is not None:
            assert rel(a.grad, b.grad) < 1e-5
    assert rel(yd.grad.float(), yr.grad.to(ydtype).float()) < 1e-6


@pytest.mark.parametrize("ydtype", [torch.float32, torch.bfloat16])
def test_unshuffle_per_sample_cls_row(ydtype):
    """y_row0 = 1: the 2D model's decoder input (OCTCube/models_mae_flash_attn.py:299-312) — row 0 of y is the sample's cls."""
    B, L, D, keep = 3, 16, 32, 4
    g = torch.Generator().manual_seed(2)
    y = torch.randn(B, 1 + keep, D, generator=g).to(ydtype)
    ids_restore = torch.stack([torch.randperm(L, generator=g) for _ in range(B)])
    mt = torch.randn(D, generator=g, requires_grad=True)
    pos = torch.randn(1 + L, D, generator=g)
    yr = y.float().clone().requires_grad_(True)
    x_ = torch.cat([yr[:, 1:], mt.expand(B, L - keep, D)], 1)
    x_ = torch.gather(x_, 1, ids_restore[..., None].expand(-1, -1, D))
    want = torch.cat([yr[:, :1], x_], 1) + pos
    dout = torch.randn(B, L + 1, D, generator=g)
    want.backward(dout)
    yd = y.detach().to(DEV).requires_grad_(True)
    mtd = mt.detach().to(DEV).requires_grad_(True)
    got = ops.UnshuffleFn.apply(yd, ids_restore.to(DEV), mtd, pos[1:].contiguous().to(DEV), None, pos[0].contiguous().to(DEV), 1)
    assert rel(got, want.detach()) < 1e-6
    got.backward(dout.to(DEV))
    assert rel(mtd.grad, mt.grad) < 1e-5
    assert rel(yd.grad.float(), yr.grad.to(ydtype).float()) < 1e-6


# ---------------------------------------------------------------- add + LayerNorm
@pytest.mark.parametrize("C", [32, 64, 512, 1024, 1280])
@pytest.mark.parametrize("hdtype,ydtype", [(torch.float32, torch.float32), (torch.bfloat16, torch.bfloat16), (torch.float32, torch.bfloat16)])
def test_add_ln_fwd_bwd(C, hdtype, ydtype):
    M = 77
    g = torch.Generator().manual_seed(C)
    h = (torch.randn(M, C, generator=g) * 2 + 0.5).to(hdtype)
    res = torch.randn(M, C, generator=g)
    gamma, beta = torch.randn(C, generator=g), torch.randn(C, generator=g)
    hr, rr = h.float().clone().requires_grad_(True), res.clone().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    r_ref = hr + rr
    y_ref = F.layer_norm(r_ref, (C,), gr, br, 1e-6)
    dy, dres = torch.randn(M, C, generator=g).to(ydtype), torch.randn(M, C, generator=g)
    (y_ref * dy.float()).sum().backward(retain_graph=True)
    r_ref.backward(dres)
    hd, rd = h.detach().to(DEV).requires_grad_(True), res.to(DEV).requires_grad_(True)
    gd, bd = gamma.to(DEV).requires_grad_(True), beta.to(DEV).requires_grad_(True)
    y, r = ops.AddLNFn.apply(hd, rd, gd, bd, 1e-6, ydtype, True)
    tol = 1e-5 if ydtype == torch.float32 else 6e-3
    assert rel(y, y_ref.detach()) < tol and rel(r, r_ref.detach()) < 1e-6
    torch.autograd.backward([y, r], [dy.to(DEV), dres.to(DEV)])
    btol = 1e-4 if hdtype == torch.float32 else 6e-3
    assert rel(rd.grad, rr.grad) < 1e-4 and rel(hd.grad, hr.grad) < btol
    assert rel(gd.grad, gr.grad) < 1e-4 and rel(bd.grad, br.grad) < 1e-4


@pytest.mark.parametrize("C", [128, 256, 512, 1024])
def test_add_ln_bwd_pipelined_many_rows(C):
    """Several rows per warp team: exercises the one-row-ahead prefetch and the two-warps-per-row statistics exchange."""
    M = 5003
    g = torch.Generator().manual_seed(C + 1)
    h = (torch.randn(M, C, generator=g) * 2 + 0.5).bfloat16()
    res = torch.randn(M, C, generator=g)
    gamma, beta = torch.randn(C, generator=g), torch.randn(C, generator=g)
    hr, rr = h.float().clone().requires_grad_(True), res.clone().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    r_ref = hr + rr
    y_ref = F.layer_norm(r_ref, (C,), gr, br, 1e-6)
    dy, dres = torch.randn(M, C, generator=g).bfloat16(), torch.randn(M, C, generator=g)
    (y_ref * dy.float()).sum().backward(retain_graph=True)
    r_ref.backward(dres)
    hd, rd = h.to(DEV).requires_grad_(True), res.to(DEV).requires_grad_(True)
    gd, bd = gamma.to(DEV).requires_grad_(True), beta.to(DEV).requires_grad_(True)
    y, r = ops.AddLNFn.apply(hd, rd, gd, bd, 1e-6, torch.bfloat16, True)
    torch.autograd.backward([y, r], [dy.to(DEV), dres.to(DEV)])
    assert rel(rd.grad, rr.grad) < 1e-4 and rel(hd.grad, hr.grad) < 6e-3
    assert rel(gd.grad, gr.grad) < 1e-4 and rel(bd.grad, br.grad) < 1e-4


def test_ln_only_final_norm_form():
    M, C = 50, 64
    h = torch.randn(M, C).bfloat16()
    gamma, beta = torch.randn(C), torch.randn(C)
    y, r = ops.AddLNFn.apply(h.to(DEV), None, gamma.to(DEV), beta.to(DEV), 1e-6, torch.bfloat16, False)
    assert r is None
    assert rel(y, F.layer_norm(h.float(), (C,), gamma, beta, 1e-6)) < 6e-3


# ---------------------------------------------------------------- GELU / colsum / cast
def test_gelu_colsum_cast():
    x = torch.randn(64, 256) * 3
    xr = x.clone().requires_grad_(True)
    F.gelu(xr).sum().backward()
    assert rel(ops.gelu_fwd(x.to(DEV)), F.gelu(x)) < 1e-6
    assert rel(ops.gelu_bwd(torch.ones_like(x).to(DEV), x.to(DEV)), xr.grad) < 1e-5
    big = torch.randn(5000, 136)
    assert rel(ops.colsum(big.to(DEV)), big.double().sum(0)) < 1e-5
    assert rel(ops.colsum(big.bfloat16().to(DEV)), big.bfloat16().double().sum(0)) < 1e-5
    assert torch.equal(ops.cast_bf16(big.to(DEV)).cpu(), big.bfloat16())
    odd = torch.randn(1027)
    assert torch.equal(ops.cast_bf16(odd.to(DEV)).cpu(), odd.bfloat16())


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_colsum_unaligned_width(dtype):
    """classifier heads: N not a multiple of 4 takes the scalar path of oct_colsum."""
    x = torch.randn(37, 5, generator=torch.Generator().manual_seed(0)).to(dtype)
    got = ops.colsum(x.to(DEV))
    assert rel(got, x.float().sum(0)) < 1e-6


# ---------------------------------------------------------------- volume ingest
def _cpu_ingest(cube, T, flip_t, flip_w):
    """The reference's CPU pipeline: ToTensor /255 (PatientDataset_inhouse.py:420), centre pad / crop of frames (:436-450),
    RandFlipd along frames / width (create_3d_transforms :59-62)."""
    out = []
    for b in range(cube.shape[0]):
        fr = cube[b].float().div(255)                      # [T_src, H, W]
        n = fr.shape[0]
        if n < T:
            left = (T - n) // 2
            fr = torch.cat([torch.zeros(left, *fr.shape[1:]), fr, torch.zeros(T - n - left, *fr.shape[1:])], 0)
        elif n > T:
            left = (n - T) // 2
            fr = fr[left:-(n - T - left)]
        if flip_t is not None and flip_t[b]:
            fr = fr.flip(0)
        if flip_w is not None and flip_w[b]:
            fr = fr.flip(2)
        out.append(fr)
    return torch.stack(out).unsqueeze(1)


@pytest.mark.parametrize("B,T_src,T,H,W", [(2, 12, 12, 64, 64), (3, 49, 60, 32, 256), (2, 61, 48, 16, 512), (1, 128, 60, 8, 36)])
def test_ingest_u8_bit_exact(B, T_src, T, H, W):
    g = torch.Generator().manual_seed(T_src)
    cube = torch.randint(0, 256, (B, T_src, H, W), generator=g, dtype=torch.uint8)
    flip_t = torch.randint(0, 2, (B,), generator=g, dtype=torch.uint8)
    flip_w = 1 - flip_t if B > 1 else torch.ones(1, dtype=torch.uint8)
    got = ops.ingest_u8(cube.to(DEV), T, flip_t=flip_t.to(DEV), flip_w=flip_w.to(DEV))
    assert torch.equal(got.cpu(), _cpu_ingest(cube, T, flip_t, flip_w))
    assert torch.equal(ops.ingest_u8(cube.to(DEV), T).cpu(), _cpu_ingest(cube, T, None, None))     # no flips


def _cpu_crop_resize(cube, T_pad, T, H, W, crop, flip_t, flip_w, device="cpu"):
    """The loader's CPU pipeline (PatientDataset_inhouse.py:420-450 + create_3d_transforms :56-63) restated with torch:
    ToTensor, centre pad / crop of frames, monai CropForegroundd (box of x > 0, margin 0), Resized = F.interpolate trilinear."""
    fr = cube.float() / 255.0
    n = fr.shape[0]
    if n < T_pad:
        left = (T_pad - n) // 2
        fr = torch.cat([torch.zeros(left, *fr.shape[1:]), fr, torch.zeros(T_pad - n - left, *fr.shape[1:])], 0)
    elif n > T_pad:
        left = (n - T_pad) // 2
        fr = fr[left:left + T_pad]
    if crop and bool((fr > 0).any()):
        nz = fr > 0
        idx = [nz.any(dim=tuple(d for d in range(3) if d != a)).nonzero().flatten() for a in range(3)]
        fr = fr[idx[0][0]:idx[0][-1] + 1, idx[1][0]:idx[1][-1] + 1, idx[2][0]:idx[2][-1] + 1]
    out = F.interpolate(fr[None, None].to(device), size=(T, H, W), mode="trilinear", align_corners=False)[0, 0]
    if flip_t:
        out = out.flip(0)
    if flip_w:
        out = out.flip(2)
    return out


@pytest.mark.parametrize("T_src,Hs,Ws,T_pad,T,H,W,crop", [(49, 96, 128, 60, 60, 64, 64, True), (61, 64, 256, 48, 48, 32, 96, True),
                                                          (25, 40, 64, 25, 60, 256, 256, False), (30, 48, 64, 40, 20, 24, 32, True)])
def test_crop_resize_trilinear_vs_interpolate(T_src, Hs, Ws, T_pad, T, H, W, crop):
    """CropForegroundd + Resized(trilinear) + flips on the device from the uint8 cube vs F.interpolate on the CPU (what the
    reference's loader runs) and on the GPU (the blend order the kernel follows).  fp32 blends of 8 corners: a few ulp."""
    g = torch.Generator().manual_seed(T_src)
    cube = torch.randint(0, 256, (T_src, Hs, Ws), generator=g, dtype=torch.uint8)
    cube[:3] = 0; cube[-2:] = 0; cube[:, :5] = 0; cube[:, -7:] = 0; cube[:, :, :9] = 0; cube[:, :, -4:] = 0   # a foreground box
    for ft, fw in ((False, False), (True, True)):
        got = ops.crop_resize_u8(cube.to(DEV), T_pad, T, H, W, crop_foreground=crop, flip_t=ft, flip_w=fw).cpu()
        ref_cpu = _cpu_crop_resize(cube, T_pad, T, H, W, crop, ft, fw)
        ref_gpu = _cpu_crop_resize(cube, T_pad, T, H, W, crop, ft, fw, device=DEV).cpu()
        assert got.shape == ref_cpu.shape
        assert float((got - ref_cpu).abs().max()) < 1e-6 and float((got - ref_gpu).abs().max()) < 1e-6
        print("bit-equal to F.interpolate on CUDA:", bool(torch.equal(got, ref_gpu)), " on CPU:", bool(torch.equal(got, ref_cpu)))
    empty = torch.zeros(T_src, Hs, Ws, dtype=torch.uint8)                   # no foreground: the whole (padded) cube is resampled
    assert float(ops.crop_resize_u8(empty.to(DEV), T_pad, T, H, W).abs().max()) == 0.0


# ---------------------------------------------------------------- loss
@pytest.mark.parametrize("norm_pix", [False, True])
@pytest.mark.parametrize("pdtype", [torch.float32, torch.bfloat16])
def test_mse_loss_fwd_bwd(norm_pix, pdtype):
    cfg = O.MAEConfig(input_size=64, num_frames=12, pred_t_dim=12, norm_pix_loss=norm_pix)
    B, L, P = 2, 4 * 16, 768
    g = torch.Generator().manual_seed(5)
    imgs = O.synthetic_volume(B, 12, 64, 64, seed=2, zero_pad_frames=1)
    pred_full = torch.randn(B, L + 1, P, generator=g).to(pdtype)
    mask = (torch.rand(B, L, generator=g) > 0.2).float()
    pr = pred_full.float().clone().requires_grad_(True)
    loss_ref, fl_ref = O.forward_loss(cfg, imgs, pr[:, 1:], mask, frame_loss=True)
    (loss_ref * 1.7).backward()
    pd = pred_full.detach().to(DEV).requires_grad_(True)
    loss, fl, _ = ops.MaskedMSELossFn.apply(imgs.to(DEV), pd, mask.to(DEV), 16, 3, 1, norm_pix, None)
    assert abs(float(loss) - float(loss_ref)) < 2e-6 * abs(float(loss_ref))
    assert rel(fl, fl_ref.detach()) < 1e-5
    (loss * 1.7).backward()
    assert rel(pd.grad.float(), pr.grad) < (1e-5 if pdtype == torch.float32 else 5e-3)
    assert float(pd.grad[:, 0].abs().max()) == 0.0


@pytest.mark.parametrize("norm_pix", [False, True])
@pytest.mark.parametrize("pdtype", [torch.float32, torch.bfloat16])
def test_mse_loss_channel_last_2d(norm_pix, pdtype):
    """The 2D model's loss (OCTCube/models_mae_flash_attn.py:331-350): (p, q, c) patch order, frame_loss over ALL patches."""
    from octcubem_b200._lib import LOSS_ALL_TOKENS, LOSS_CHANNEL_LAST
    from oracle import mae2d_oracle as O2
    cfg = O2.MAE2DConfig(input_size=64, norm_pix_loss=norm_pix)
    B, L, P = 2, 16, 768
    g = torch.Generator().manual_seed(6)
    imgs = O2.synthetic_images(B, 3, 64, 64, seed=3)
    pred_full = torch.randn(B, L + 1, P, generator=g).to(pdtype)
    mask = (torch.rand(B, L, generator=g) > 0.3).float()
    pr = pred_full.float().clone().requires_grad_(True)
    loss_ref, fl_ref = O2.forward_loss(cfg, imgs, pr[:, 1:], mask, return_frame_loss=True)
    (loss_ref * 0.6).backward()
    pd = pred_full.detach().to(DEV).requires_grad_(True)
    vol = imgs.view(B, 1, 3, 64, 64).to(DEV)
    loss, _, tok = ops.MaskedMSELossFn.apply(vol, pd, mask.to(DEV), 16, 3, 1, norm_pix, None, LOSS_CHANNEL_LAST | LOSS_ALL_TOKENS)
    assert abs(float(loss) - float(loss_ref)) < 2e-6 * abs(float(loss_ref))
    assert rel(tok.mean(-1), fl_ref.detach()) < 1e-5
    (loss * 0.6).backward()
    assert rel(pd.grad.float(), pr.grad) < (1e-5 if pdtype == torch.float32 else 5e-3)
    loss_m, _, tok_m = ops.MaskedMSELossFn.apply(vol, pd.detach(), mask.to(DEV), 16, 3, 1, norm_pix, None, LOSS_CHANNEL_LAST)
    assert float(loss_m) == float(loss) and bool((tok_m[mask.to(DEV) == 0] == 0).all())   # masked-only mode: same loss


def test_mse_loss_many_frames_and_trailing_rows():
    """More than 4096 (b, t') groups (the finish kernel's staging tile) and a prediction buffer with rows BEHIND the L tokens
    (their gradient rows must be zero, their mask must not be read)."""
    cfg = O.MAEConfig(input_size=8, patch_size=4, num_frames=16, pred_t_dim=16, t_patch_size=1)
    B, L, P = 300, 16 * 4, 16                                                     # B * T' = 4800 groups of 4 tokens
    g = torch.Generator().manual_seed(8)
    imgs = O.synthetic_volume(B, 16, 8, 8, seed=5, zero_pad_frames=0)
    pred_full = torch.randn(B, 1 + L + 3, P, generator=g)
    mask = (torch.rand(B, L, generator=g) > 0.4).float()
    pr = pred_full.clone().requires_grad_(True)
    loss_ref, fl_ref = O.forward_loss(cfg, imgs, pr[:, 1:1 + L], mask, frame_loss=True)
    loss_ref.backward()
    pd = pred_full.to(DEV).requires_grad_(True)
    loss, fl, _ = ops.MaskedMSELossFn.apply(imgs.to(DEV), pd, mask.to(DEV), 4, 1, 1, False, None)
    assert abs(float(loss) - float(loss_ref)) < 5e-6 * abs(float(loss_ref))
    assert rel(fl, fl_ref.detach()) < 1e-5
    loss.backward()
    assert rel(pd.grad, pr.grad) < 1e-5
    assert float(pd.grad[:, 1 + L:].abs().max()) == 0.0 and float(pd.grad[:, 0].abs().max()) == 0.0


def test_mse_loss_frame_index_select():
    # pred_t_dim != T: linspace index_select of models...:630-640
    cfg = O.MAEConfig(input_size=64, num_frames=12, pred_t_dim=6, t_patch_size=2)  # u = 1, T' = 6
    imgs = O.synthetic_volume(1, 12, 64, 64, seed=4, zero_pad_frames=0)
    L, P = 6 * 16, 256
    pred = torch.randn(1, L, P)
    mask = torch.ones(1, L)
    want = O.forward_loss(cfg, imgs, pred, mask)
    fidx = torch.linspace(0, 11, 6).long()
    loss, _, _ = ops.MaskedMSELossFn.apply(imgs.to(DEV), pred.to(DEV), mask.to(DEV), 16, 1, 0, False, fidx.to(DEV))
    assert abs(float(loss) - float(want)) < 2e-6 * abs(float(want))


# ---------------------------------------------------------------- GEMMs / attention / patch-embed
@pytest.mark.parametrize("layout", [GEMM_NT, GEMM_NN, GEMM_TN])
@pytest.mark.parametrize("compute,M,N,K", [(OCT_F32, 130, 72, 40), (OCT_BF16, 384, 768, 512), (OCT_BF16, 3280, 1024, 1024),
                                           (OCT_BF16, 200, 136, 72), (OCT_BF16, 4104, 512, 512), (OCT_BF16, 512, 512, 8200), (OCT_BF16, 2048, 512, 32776)])
def test_gemm(layout, compute, M, N, K):
    g = torch.Generator().manual_seed(M)
    dt = torch.float32 if compute == OCT_F32 else torch.bfloat16
    a, b = torch.randn(M, K, generator=g).to(dt), torch.randn(N, K, generator=g).to(dt)
    A = a.to(DEV) if layout != GEMM_TN else a.t().contiguous().to(DEV)
    Bm = b.to(DEV) if layout == GEMM_NT else b.t().contiguous().to(DEV)
    out = ops.gemm(layout, A, Bm, M, N, K, torch.float32, compute=compute)
    assert rel(out, a.double() @ b.double().t()) < (2e-6 if K <= 8192 else 2e-5)  # fp32 accumulation over K


@pytest.mark.parametrize("layout,M,N,K,epi", [(GEMM_NT, 3280, 1024, 1024, "bias"), (GEMM_NT, 3280, 3072, 1024, "bias"),
                                              (GEMM_NN, 3280, 1024, 4096, "none"), (GEMM_NN, 3280, 1024, 3072, "none"),
                                              (GEMM_NT, 3280, 4096, 1024, "gelu"), (GEMM_NN, 3280, 4096, 1024, "dgelu"),
                                              (GEMM_NT, 3000, 1000, 576, "bias"), (GEMM_NT, 32776, 1536, 512, "bias"),
                                              (GEMM_NN, 32776, 512, 512, "none"), (GEMM_NT, 3280, 2064, 1024, "gelu")])
def test_gemm_bf16_out_step_shapes(layout, M, N, K, epi):
    """bf16-output GEMMs at the step's shapes (pair and single-CTA tiles, ragged M / N tails) with every fused epilogue,
    against fp64."""
    from octcubem_b200._lib import EPI_BIAS, EPI_BIAS_GELU, EPI_DGELU, EPI_NONE
    g = torch.Generator().manual_seed(M + N + K)
    a = (torch.randn(M, K, generator=g) * 0.5).bfloat16()
    b = (torch.randn(N, K, generator=g) * 0.1).bfloat16()
    bias = torch.randn(N, generator=g)
    pre = torch.randn(M, N, generator=g).bfloat16()
    Bm = b.to(DEV) if layout == GEMM_NT else b.t().contiguous().to(DEV)
    acc = a.double() @ b.double().t()
    code = {"none": EPI_NONE, "bias": EPI_BIAS, "gelu": EPI_BIAS_GELU, "dgelu": EPI_DGELU}[epi]
    aux = None
    if epi == "gelu":
        aux = torch.empty(M, N, dtype=torch.bfloat16, device=DEV)
    elif epi == "dgelu":
        aux = pre.to(DEV)
    out = ops.gemm(layout, a.to(DEV), Bm, M, N, K, torch.bfloat16, code, bias=bias.to(DEV) if epi in ("bias", "gelu") else None,
                   aux=aux, compute=OCT_BF16)
    if epi == "none":
        want = acc
    elif epi == "bias":
        want = acc + bias.double()
    elif epi == "gelu":
        z = (acc + bias.double()).to(torch.bfloat16)          # GELU of the bf16-rounded pre-activation
        assert rel(aux, acc + bias.double()) < 4e-3
        want = F.gelu(z.double())
    else:
        x = pre.double().requires_grad_(True)
        F.gelu(x).backward(acc)
        want = x.grad
    assert rel(out, want) < 4e-3
    assert torch.isfinite(out.float()).all()


@pytest.mark.parametrize("n_out,k_in,tokens", [(3072, 1024, 3280), (1024, 1024, 3280), (1536, 512, 32776), (512, 2048, 4104),
                                               (768, 512, 8200), (200, 136, 72), (128, 128, 2000), (1024, 768, 3272)])
def test_wgrad_bias_fused(n_out, k_in, tokens):
    """oct_gemm_wgrad_bias: dW = dY^T X and db = column sums of dY from one kernel (pair / single CTA, split-K or not,
    128- and 256-wide tiles, ragged M and K tails)."""
    g = torch.Generator().manual_seed(n_out + tokens)
    dy = torch.randn(tokens, n_out, generator=g).bfloat16()
    x = torch.randn(tokens, k_in, generator=g).bfloat16()
    dw, db = ops.wgrad_bias(dy.to(DEV), x.to(DEV))
    assert dw.dtype == torch.float32 and db.dtype == torch.float32
    assert rel(dw, dy.double().t() @ x.double()) < (2e-6 if tokens <= 8192 else 2e-5)
    assert rel(db, dy.double().sum(0)) < 2e-6
    dw2, db2 = ops.wgrad_bias(dy.to(DEV), x.to(DEV))  # repeatable up to the split-K reduction order
    assert rel(dw2, dw) < 1e-6 and rel(db2, db) < 1e-6


def test_linear_and_mlp_functions_bf16():
    g = torch.Generator().manual_seed(9)
    M, dim, hid = 300, 128, 512
    x = torch.randn(M, dim, generator=g)
    w1, b1 = torch.randn(hid, dim, generator=g) * 0.1, torch.randn(hid, generator=g) * 0.1
    w2, b2 = torch.randn(dim, hid, generator=g) * 0.1, torch.randn(dim, generator=g) * 0.1
    leaves = [t.clone().requires_grad_(True) for t in (x, w1, b1, w2, b2)]
    y_ref = F.linear(F.gelu(F.linear(leaves[0], leaves[1], leaves[2])), leaves[3], leaves[4])
    dy = torch.randn(M, dim, generator=g)
    y_ref.backward(dy)
    d = [t.to(DEV).requires_grad_(True) for t in (x.bfloat16(), w1, b1, w2, b2)]
    y = ops.MlpFn.apply(d[0], d[1], d[2], d[3], d[4], d[1].detach().bfloat16(), d[3].detach().bfloat16())
    assert rel(y.float(), y_ref.detach()) < 1e-2
    y.backward(dy.bfloat16().to(DEV))
    for got, want in zip(d, leaves):
        assert rel(got.grad.float(), want.grad) < 2e-2


@pytest.mark.parametrize("compute,dtype,tol", [(OCT_F32, torch.float32, 2e-5), (OCT_BF16, torch.bfloat16, 8e-3)])
# 1030 / 1060: the last 256-row CTA holds a single live Q tile and sweeps 9 kv tiles (run-ahead 2); 1160: two live Q tiles
# and a short last kv tile; 1030 / 1160 also have an odd number of 64-query sub-tiles in the backward sweep
# 4097 / 5121: the production decoder lengths of cfg-1/2 and cfg-3 (32*128+1 / 40*128+1: the single-live-row tail CTA);
# (8,410,16,64) / (2,512,16,64): the encoder shapes; 385 / 449 = 3*128+1 / 3*128+65 per head dim (tail sub-tiles)
@pytest.mark.parametrize("B,S,H,d", [(2, 77, 2, 32), (1, 300, 2, 64), (2, 512, 4, 64), (1, 1030, 3, 32), (1, 1060, 2, 32),
                                     (1, 1160, 2, 32), (1, 1030, 1, 64), (1, 4097, 2, 32), (1, 5121, 2, 32), (8, 410, 16, 64),
                                     (2, 512, 16, 64), (1, 385, 2, 32), (1, 449, 2, 32), (1, 385, 2, 64), (1, 449, 2, 64)])
def test_attention(compute, dtype, tol, B, S, H, d):
    import math
    g = torch.Generator().manual_seed(S)
    qkv = torch.randn(B, S, 3 * H * d, generator=g).to(dtype)
    dout = torch.randn(B, S, H * d, generator=g).to(dtype)
    x = qkv.double().requires_grad_(True)
    q, k, v = x.view(B, S, 3, H, d).unbind(2)
    s = torch.einsum("bthd,bshd->bhts", q, k) / math.sqrt(d)
    o_ref = torch.einsum("bhts,bshd->bthd", torch.softmax(s, -1), v).reshape(B, S, H * d)
    (o_ref * dout.double()).sum().backward()
    qd = qkv.to(DEV).requires_grad_(True)
    out = ops.AttnFn.apply(qd, H, compute)
    assert rel(out, o_ref.detach()) < tol
    out.backward(dout.to(DEV))
    assert rel(qd.grad, x.grad) < 2 * tol


@pytest.mark.parametrize("S", [4097, 5121])
def test_attention_at_the_bench_shape(S):
    """The exact launch that is 47 % of the step: B = 8, H = 16, d = 32, S = 4097 (cfg-1/2) and 5121 (cfg-3), forward and
    backward, checked on a spread of (batch, head) pairs against an fp64 evaluation of the same bf16 inputs (torch on the
    GPU is only the checker here)."""
    import math
    B, H, d = 8, 16, 32
    g = torch.Generator().manual_seed(S)
    qkv = (torch.randn(B, S, 3 * H * d, generator=g) * 0.7).bfloat16().to(DEV).requires_grad_(True)
    dout = torch.randn(B, S, H * d, generator=g).bfloat16().to(DEV)
    out = ops.AttnFn.apply(qkv, H, OCT_BF16)
    out.backward(dout)
    q5 = qkv.detach().view(B, S, 3, H, d)
    worst_o = worst_g = 0.0
    for b, h in ((0, 0), (3, 7), (7, 15), (5, 1)):
        x = q5[b, :, :, h, :].double().clone().requires_grad_(True)          # [S, 3, d]
        q, k, v = x.unbind(1)
        p = torch.softmax(q @ k.t() / math.sqrt(d), -1)
        o_ref = p @ v
        (o_ref * dout[b, :, h * d:(h + 1) * d].double()).sum().backward()
        worst_o = max(worst_o, rel(out[b, :, h * d:(h + 1) * d], o_ref.detach()))
        worst_g = max(worst_g, rel(qkv.grad.view(B, S, 3, H, d)[b, :, :, h, :], x.grad))
    print(f"S={S}: worst out rel {worst_o:.2e}, worst dqkv rel {worst_g:.2e}")
    assert worst_o < 8e-3 and worst_g < 1.6e-2
    # the tail row (query S-1 lives alone in the last 128-row tile) and the cls row
    assert torch.isfinite(out).all() and torch.isfinite(qkv.grad).all()


@pytest.mark.parametrize("B,T,HW,E,u", [(2, 6, 64, 64, 3), (1, 12, 256, 1024, 3), (2, 3, 128, 264, 3)])
def test_patch_embed_tc(B, T, HW, E, u):
    g = torch.Generator().manual_seed(E)
    imgs = torch.rand(B, 1, T, HW, HW, generator=g)
    w, b = torch.randn(E, 1, u, 16, 16, generator=g) * 0.05, torch.randn(E, generator=g)
    want = O.patch_embed(imgs.double(), w.double(), b.double()).reshape(B, -1, E)
    out = ops.patch_embed_tc(imgs.to(DEV), w.view(E, -1).to(DEV).contiguous(), b.to(DEV), 16, u, torch.float32)
    assert rel(out, want) < 2e-3  # tf32 operands (10-bit mantissa), fp32 accumulate


def test_cast_bf16_multi_matches_the_per_tensor_cast():
    """oct_cast_f32_to_bf16_multi (all weight shadows in one launch, chunk table) against the per-tensor cast: bit-equal,
    including tails that are not multiples of 4 and tensors longer than one chunk."""
    g = torch.Generator().manual_seed(5)
    srcs = [torch.randn(n, generator=g).to(DEV) for n in (7, 16384, 16385, 40000, 1024 * 1024 + 3, 64)]
    dsts = [torch.empty(t.numel(), dtype=torch.bfloat16, device=DEV) for t in srcs]
    table, n = ops.cast_table(list(zip(srcs, dsts)))
    assert n == sum((t.numel() + ops.CAST_CHUNK - 1) // ops.CAST_CHUNK for t in srcs)
    ops.cast_bf16_multi(table, n)
    for s_, d_ in zip(srcs, dsts):
        assert torch.equal(d_, ops.cast_bf16(s_))
        assert torch.equal(d_, s_.bfloat16())


def test_gemm_tc_from_fresh_thread():
    """A thread without a bound CUDA context (a new autograd worker) must be able to call the TMA-based GEMM."""
    import threading
    a = torch.randn(256, 128, device=DEV).bfloat16()
    b = torch.randn(192, 128, device=DEV).bfloat16()
    box = {}

    def work():
        try:
            box["out"] = ops.gemm(GEMM_NT, a, b, 256, 192, 128, torch.float32, compute=OCT_BF16)
        except Exception as e:  # noqa: BLE001
            box["err"] = e

    t = threading.Thread(target=work)
    t.start()
    t.join()
    assert "err" not in box, box.get("err")
    assert rel(box["out"], a.double() @ b.double().t()) < 2e-6


# ---------------------------------------------------------------- optimizer
def test_fused_adamw_matches_torch():
    """oct_adamw_step vs torch.optim.AdamW (the reference's optimizer, main_pretrain...:451-455) over 4 steps, two groups
    (decay / no decay, misc.add_weight_decay), a changing learning rate, odd tensor sizes, and the bf16 shadow output."""
    from octcubem_b200 import optim
    torch.manual_seed(3)
    net = torch.nn.Sequential(torch.nn.Linear(37, 129), torch.nn.LayerNorm(129), torch.nn.Linear(129, 20001)).to(DEV)
    ref = torch.nn.Sequential(torch.nn.Linear(37, 129), torch.nn.LayerNorm(129), torch.nn.Linear(129, 20001)).to(DEV)
    ref.load_state_dict(net.state_dict())
    groups = optim.add_weight_decay(net, 0.05)
    assert len(groups[0]["params"]) == 4 and len(groups[1]["params"]) == 2 and groups[0]["weight_decay"] == 0.0
    shadows = {id(p): torch.zeros(p.shape, dtype=torch.bfloat16, device=DEV) for p in net.parameters() if p.dim() == 2}
    opt = optim.FusedAdamW(groups, lr=1e-3, betas=(0.9, 0.95), shadows=lambda p: shadows.get(id(p)))
    ropt = torch.optim.AdamW(optim.add_weight_decay(ref, 0.05), lr=1e-3, betas=(0.9, 0.95))
    for step in range(4):
        lr = optim.adjust_learning_rate(opt, step * 0.5, 1e-3, 1e-5, 1, 3)
        assert lr == optim.adjust_learning_rate(ropt, step * 0.5, 1e-3, 1e-5, 1, 3)
        g = torch.Generator().manual_seed(step)
        x = torch.randn(8, 37, generator=g).to(DEV)
        for m, o in ((net, opt), (ref, ropt)):
            m.zero_grad(set_to_none=True)
            (m(x) ** 2).mean().backward()
        v0 = net[0].weight._version
        opt.step()
        ropt.step()
        assert net[0].weight._version > v0
        for (k, p), (_, r) in zip(net.named_parameters(), ref.named_parameters()):
            assert rel(p.detach(), r.detach()) < 1e-6, (step, k)
    for p in net.parameters():
        if p.dim() == 2:
            assert torch.equal(shadows[id(p)], p.detach().bfloat16())


@pytest.mark.parametrize("use_graph", [False, True])
def test_fused_adamw_device_clock_follows_the_cosine_schedule(use_graph):
    """FusedAdamW(schedule=...): step count, bias corrections and the per-iteration cosine lr (lr_sched.py:10-28) live in a
    device clock, so REPLAYING one captured optimizer step performs consecutive steps — checked against torch.optim.AdamW
    driven by the host-side schedule, across warm-up and decay, with a group lr_scale and gradient clipping."""
    from octcubem_b200 import optim
    torch.manual_seed(5)
    net = torch.nn.Sequential(torch.nn.Linear(33, 65), torch.nn.LayerNorm(65), torch.nn.Linear(65, 4099)).to(DEV)
    ref = torch.nn.Sequential(torch.nn.Linear(33, 65), torch.nn.LayerNorm(65), torch.nn.Linear(65, 4099)).to(DEV)
    ref.load_state_dict(net.state_dict())
    sched = optim.CosineSchedule(lr=2e-3, min_lr=1e-5, warmup_epochs=1.0, epochs=4.0, epochs_per_step=0.5)
    groups, rgroups = optim.add_weight_decay(net, 0.05), optim.add_weight_decay(ref, 0.05)
    groups[0]["lr_scale"] = rgroups[0]["lr_scale"] = 0.5
    opt = optim.FusedAdamW(groups, lr=123.0, betas=(0.9, 0.95), schedule=sched)      # the group lr is ignored
    ropt = torch.optim.AdamW(rgroups, lr=1.0, betas=(0.9, 0.95))
    grads = [torch.zeros_like(p) for p in net.parameters()]                              # static gradient buffers
    for p, g in zip(net.parameters(), grads):
        p.grad = g
    graph = None
    for k in range(1, 8):
        gen = torch.Generator().manual_seed(k)
        fresh = [torch.randn(p.shape, generator=gen).to(DEV) * 3.0 for p in net.parameters()]
        for g, f, r in zip(grads, fresh, ref.parameters()):
            g.copy_(f)
            r.grad = f.clone()
        if not use_graph:
            opt.step(max_grad_norm=1.0)
        elif graph is None:
            opt.prepare(max_grad_norm=1.0)                                               # state / tables / clock before capture
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):                                                # capture does not execute
                opt.step(max_grad_norm=1.0)
            graph.replay()
        else:
            graph.replay()
        lr = sched.lr_at_step(k)
        assert lr == pytest.approx(optim.adjust_learning_rate(ropt, (k - 1) * 0.5, 2e-3, 1e-5, 1.0, 4.0), rel=1e-12)
        torch.nn.utils.clip_grad_norm_(ref.parameters(), 1.0)
        ropt.step()
        step, dev_lr = opt.clock_state()
        assert step == k and dev_lr == pytest.approx(lr, rel=1e-6, abs=1e-12)
        for (name, p), (_, r) in zip(net.named_parameters(), ref.named_parameters()):
            assert rel(p.detach(), r.detach()) < 2e-6, (k, name)


def test_fused_adamw_checkpoint_resume_under_graph_replay():
    """optimizer.state_dict() / load_state_dict() on the clocked (graph-replayed) path (the reference checkpoints through
    misc.save_model -> optimizer.state_dict()): the step count comes from the DEVICE clock (the host mirror does not advance
    under replay), a resumed optimizer continues the cosine schedule and the bias corrections where the checkpoint left them,
    and its kernels follow the loaded moment tensors (not the ones the tables were built for)."""
    from octcubem_b200 import optim
    sched = optim.CosineSchedule(lr=2e-3, min_lr=1e-5, warmup_epochs=1.0, epochs=4.0, epochs_per_step=0.5)

    def make():
        torch.manual_seed(11)
        net = torch.nn.Sequential(torch.nn.Linear(33, 65), torch.nn.LayerNorm(65), torch.nn.Linear(65, 515)).to(DEV)
        opt = optim.FusedAdamW(optim.add_weight_decay(net, 0.05), lr=1.0, betas=(0.9, 0.95), schedule=sched)
        grads = [torch.zeros_like(p) for p in net.parameters()]
        for p, g in zip(net.parameters(), grads):
            p.grad = g
        return net, opt, grads

    def feed(grads, k):
        gen = torch.Generator().manual_seed(100 + k)
        for g in grads:
            g.copy_(torch.randn(g.shape, generator=gen).to(DEV))

    def graph_of(opt):
        opt.prepare()
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            opt.step()
        return gr

    # uninterrupted: 6 replayed steps
    net_a, opt_a, grads_a = make()
    gr = graph_of(opt_a)
    for k in range(1, 7):
        feed(grads_a, k)
        gr.replay()
    # interrupted after 3 steps: checkpoint, fresh objects, resume
    net_b, opt_b, grads_b = make()
    gr_b = graph_of(opt_b)
    for k in range(1, 4):
        feed(grads_b, k)
        gr_b.replay()
    ck_opt, ck_net = opt_b.state_dict(), {k: v.clone() for k, v in net_b.state_dict().items()}
    assert all(int(st["step"]) == 3 for st in ck_opt["state"].values())              # read from the device clock
    net_c, opt_c, grads_c = make()
    net_c.load_state_dict(ck_net)
    opt_c.prepare()                                                                  # tables exist BEFORE the load
    opt_c.load_state_dict(ck_opt)
    assert opt_c.clock_state()[0] == 3
    gr_c = graph_of(opt_c)
    for k in range(4, 7):
        feed(grads_c, k)
        gr_c.replay()
    assert opt_c.clock_state()[0] == 6 and opt_c.clock_state()[1] == pytest.approx(sched.lr_at_step(6), rel=1e-6)
    for (name, p), (_, q) in zip(net_a.named_parameters(), net_c.named_parameters()):
        assert torch.equal(p.detach(), q.detach()), name
    for pa, pc in zip(net_a.parameters(), net_c.parameters()):
        assert torch.equal(opt_a.state[pa]["exp_avg_sq"], opt_c.state[pc]["exp_avg_sq"])


def test_fused_adamw_grad_scale_and_errors():
    from octcubem_b200 import optim
    p = torch.nn.Parameter(torch.randn(1000, device=DEV))
    q = torch.nn.Parameter(p.detach().clone())
    g = torch.randn(1000, device=DEV)
    p.grad, q.grad = g * 1024.0, g.clone()
    a, b = optim.FusedAdamW([p], lr=1e-2), optim.FusedAdamW([q], lr=1e-2)
    a.step(grad_scale=1.0 / 1024.0)   # GradScaler unscale folded into the update
    b.step()
    assert rel(p.detach(), q.detach()) < 1e-6
    cpu = torch.nn.Parameter(torch.randn(4))
    cpu.grad = torch.randn(4)
    with pytest.raises(RuntimeError):
        optim.FusedAdamW([cpu]).step()


def test_fused_adamw_clip_grad_norm():
    """Device-side global norm + clip coefficient (oct_grad_norm) folded into the update, vs clip_grad_norm_ + AdamW."""
    from octcubem_b200 import optim
    torch.manual_seed(5)
    net = torch.nn.Sequential(torch.nn.Linear(33, 257), torch.nn.Linear(257, 9)).to(DEV)
    ref = torch.nn.Sequential(torch.nn.Linear(33, 257), torch.nn.Linear(257, 9)).to(DEV)
    ref.load_state_dict(net.state_dict())
    opt = optim.FusedAdamW(optim.add_weight_decay(net, 0.05), lr=1e-3, betas=(0.9, 0.95))
    ropt = torch.optim.AdamW(optim.add_weight_decay(ref, 0.05), lr=1e-3, betas=(0.9, 0.95))
    for step, max_norm in enumerate((0.05, 1e6, 0.3)):   # clipping, not clipping, clipping
        x = torch.randn(16, 33, generator=torch.Generator().manual_seed(step)).to(DEV)
        for m in (net, ref):
            m.zero_grad(set_to_none=True)
            (m(x) ** 2).mean().backward()
        want = torch.nn.utils.clip_grad_norm_(ref.parameters(), max_norm)
        opt.step(max_grad_norm=max_norm)
        ropt.step()
        assert abs(float(opt.grad_norm) - float(want)) < 1e-5 * float(want)
        for (k, p), (_, r) in zip(net.named_parameters(), ref.named_parameters()):
            assert rel(p.detach(), r.detach()) < 2e-6, (step, k)
