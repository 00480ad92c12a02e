"""GPU parity of the 2D twin (OCTCube/models_mae_flash_attn.py surface -> C ABI -> the same kernels as the 3D model)
against the committed reference-generated fixture tests/golden/toy2d_step.npz and the CPU oracle."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from octcubem_b200 import models_mae_flash_attn as M2  # noqa: E402
from oracle import mae2d_oracle as O2  # noqa: E402
from oracle import mae3d_oracle as O  # noqa: E402
from oracle.gen_golden import TOY2D, toy2d_inputs  # noqa: E402

DEV = "cuda:0"
FP32_TOL = 1e-4   # north star: loss and gradients within 1e-4 relative in fp32
BF16_TOL = 2e-2   # ... and 2e-2 relative in bf16


def rel(a, b):
    a, b = a.double().cpu(), torch.as_tensor(b).double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def build(cfg, sd, precision):
    m = M2.MaskedAutoencoderViT(**cfg.ref_kwargs(), precision=precision,
                                norm_layer=lambda d: torch.nn.LayerNorm(d, eps=cfg.ln_eps)).to(DEV)
    m.load_state_dict(sd, strict=True)
    return m


@pytest.mark.parametrize("precision,tol", [("fp32", FP32_TOL), ("bf16", BF16_TOL)])
def test_toy2d_step_vs_reference_golden(golden_dir, precision, tol):
    g = np.load(os.path.join(golden_dir, "toy2d_step.npz"))
    sd, imgs, noise = toy2d_inputs()
    m = build(TOY2D, sd, precision)
    loss, pred, mask, frame_loss = m(imgs.to(DEV), mask_ratio=0.75, return_frame_loss=True, noise=noise.to(DEV))
    loss.backward()
    assert np.array_equal(mask.cpu().numpy(), g["mask"])                      # bit-exact
    assert abs(float(loss) - float(g["loss"])) < tol * abs(float(g["loss"]))
    assert rel(frame_loss, g["frame_loss"]) < tol
    assert rel(pred.float(), g["pred"]) < tol
    grads = {k: p.grad for k, p in m.named_parameters()}
    assert grads["pos_embed"] is None and grads["decoder_pos_embed"] is None  # frozen sin-cos tables (:97,143)
    worst = 0.0
    for k in g.files:
        if k.startswith("g::"):
            r = rel(grads[k[3:]].float(), g[k])
            worst = max(worst, r)
            assert r < (tol), (k, r)
    print(f"2D worst grad rel err ({precision}): {worst:.2e}")


def test_toy2d_step_normpix_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "toy2d_step.npz"))
    sd, imgs, noise = toy2d_inputs()
    cfg = O2.MAE2DConfig(**{**TOY2D.__dict__, "norm_pix_loss": True})
    m = build(cfg, sd, "fp32")
    loss, pred, mask = m(imgs.to(DEV), mask_ratio=0.75, noise=noise.to(DEV))
    loss.backward()
    assert abs(float(loss) - float(g["np::loss"])) < FP32_TOL * abs(float(g["np::loss"]))
    params = dict(m.named_parameters())
    for k in g.files:
        if k.startswith("np::g::"):
            assert rel(params[k[7:]].grad, g[k]) < FP32_TOL, k


@pytest.mark.parametrize("precision,tol", [("fp32", FP32_TOL), ("bf16", BF16_TOL)])
def test_2d_larger_shape_vs_oracle(precision, tol):
    """128-px input (64 patches, several attention tiles' worth of decoder rows at B=3), mask 0.85 -> keep = 9."""
    cfg = O2.MAE2DConfig(**{**TOY2D.__dict__, "input_size": 128})
    sd = O.perturb_state_dict(O2.init_state_dict(cfg, seed=2))
    imgs = O2.synthetic_images(3, 3, 128, 128, seed=4)
    noise = O.synthetic_noise(3, 64, seed=8)
    (ref_loss, ref_pred, ref_mask, ref_fl), ref_g = O2.forward_backward(cfg, sd, imgs, 0.85, noise)
    m = build(cfg, sd, precision)
    loss, pred, mask, fl = m(imgs.to(DEV), mask_ratio=0.85, return_frame_loss=True, noise=noise.to(DEV))
    loss.backward()
    assert torch.equal(mask.cpu(), ref_mask) and float(mask.sum()) == 3 * (64 - 9)
    assert abs(float(loss) - float(ref_loss)) < tol * abs(float(ref_loss))
    assert rel(fl, ref_fl.detach()) < tol
    got = {k: p.grad for k, p in m.named_parameters() if p.grad is not None}
    assert set(got) == set(ref_g)
    for k in got:
        assert rel(got[k], ref_g[k]) < (tol), k


def test_2d_module_surface_and_methods():
    sd, imgs, noise = toy2d_inputs()
    m = build(TOY2D, sd, "fp32")
    x = m.patch_embed(imgs.to(DEV))
    w = sd["patch_embed.proj.weight"]
    want = torch.nn.functional.conv2d(imgs, w, sd["patch_embed.proj.bias"], stride=16).flatten(2).transpose(1, 2)
    assert rel(x, want) < 1e-5
    out = m.random_masking(x, 0.75, noise=noise.to(DEV))
    assert len(out) == 3                                                       # 3-tuple, unlike the 3D model (:267)
    o = O2.random_masking(want, 0.75, noise)
    assert torch.equal(out[1].cpu(), o[1]) and torch.equal(out[2].cpu(), o[2])
    assert torch.equal(out[0].cpu(), torch.gather(x.cpu(), 1, o[3][..., None].expand(-1, -1, 64)))   # exact row copy
    # public forward_encoder / forward_decoder / forward_loss chain == forward()
    latent, mask, ids_restore = m.forward_encoder(imgs.to(DEV), 0.75, noise=noise.to(DEV))
    assert latent.shape == (2, 1 + 4, 64)                                     # the cls token is kept (:296)
    pred = m.forward_decoder(latent, ids_restore)
    loss = m.forward_loss(imgs.to(DEV), pred, mask)
    loss2, pred2, mask2 = m(imgs.to(DEV), mask_ratio=0.75, noise=noise.to(DEV))
    assert float(loss) == float(loss2) and torch.equal(pred, pred2) and torch.equal(mask, mask2)
    p = m.patchify(imgs.to(DEV))
    assert torch.equal(p.cpu(), O2.patchify(imgs, 16)) and torch.equal(m.unpatchify(p).cpu(), imgs)
    with pytest.raises(AssertionError):
        m(torch.zeros(1, 3, 32, 32, device=DEV))                              # :63-66
