"""CPU tests of the oracle: against the committed reference-generated fixtures, and (where the reference
tree is present, i.e. in the build container) against the unmodified reference itself."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import mae2d_oracle as O2
from oracle import mae3d_oracle as O
from oracle import ref_harness as R
from oracle import vit_st_oracle as OV
from oracle.gen_golden import TOY, TOY2D, TOY_VIT, toy2d_inputs, toy_inputs, toy_vit_inputs

needs_ref = pytest.mark.skipif(not R.reference_available(), reason="/root/reference not present")


def _rel(a, b):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    return float((a - b).norm() / (b.norm() + 1e-30))


def test_len_keep_truncation():
    # quirk Q2 (models_mae_joint_res_flash_attn.py:349)
    assert O.len_keep_of(5120, 0.9) == 511
    assert O.len_keep_of(4096, 0.9) == 409
    assert O.len_keep_of(1024, 0.75) == 256
    assert O.len_keep_of(1024, 0.85) == 153


def test_patchify_roundtrip_and_gemm_view():
    x = torch.randn(2, 1, 12, 64, 64)
    p = O.patchify(x, 16, 3)
    assert p.shape == (2, 4 * 16, 768)
    assert torch.equal(O.unpatchify(p, 2, 1, 12, 64, 64, 16, 3), x)
    # patchify(imgs) is exactly the A matrix of the patch-embed GEMM for C=1
    w, b = torch.randn(32, 1, 3, 16, 16), torch.randn(32)
    y = O.patch_embed(x, w, b).reshape(2, 64, 32)
    assert torch.allclose(y, p @ w.view(32, -1).t() + b, atol=1e-4, rtol=1e-4)


def test_masking_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "masking_cases.npz"))
    for L in (1024, 4096, 5120):
        for kind in ("natural", "tiefree", "quantized"):
            noise = torch.from_numpy(g[f"{kind}_{L}_noise"])
            x = torch.arange(2 * L * 2, dtype=torch.float32).view(2, L, 2)
            xm, mask, ids_restore, ids_keep = O.random_masking(x, 0.9, noise)
            assert np.array_equal(ids_restore.numpy(), g[f"{kind}_{L}_ids_restore"].astype(np.int64)), (kind, L)
            assert np.array_equal(ids_keep.numpy(), g[f"{kind}_{L}_ids_keep"].astype(np.int64))
            assert np.array_equal(mask.numpy().astype(np.uint8), g[f"{kind}_{L}_mask"])
            # equivalences the CUDA path relies on (SURVEY §8a random_masking row)
            keep = ids_keep.shape[1]
            assert torch.equal(mask, (ids_restore >= keep).float())
            assert torch.equal(xm, torch.gather(x, 1, ids_keep[..., None].expand(-1, -1, 2)))


@pytest.mark.parametrize("norm_pix", [False, True])
def test_toy_step_golden(golden_dir, norm_pix):
    g = np.load(os.path.join(golden_dir, "toy_step_normpix.npz" if norm_pix else "toy_step.npz"))
    cfg = O.MAEConfig(**{**TOY.__dict__, "norm_pix_loss": norm_pix})
    sd, vol, noise = toy_inputs()
    assert np.array_equal(vol.numpy(), g["volume"]) and np.array_equal(noise.numpy(), g["noise"])
    (out, grads) = O.forward_backward(cfg, sd, vol, 0.9, noise, frame_loss=True)
    (loss, fl), pred, mask = out
    assert abs(float(loss) - float(g["loss"])) <= 1e-6 * abs(float(g["loss"]))
    assert np.array_equal(mask.numpy(), g["mask"])
    assert _rel(fl.detach(), g["frame_losses"]) < 1e-6
    if not norm_pix:
        assert _rel(pred.detach(), g["pred"]) < 1e-6
    n = 0
    for k in g.files:
        if k.startswith("g::"):
            assert _rel(grads[k[3:]], g[k]) < 2e-5, k
            n += 1
    assert n >= 4
    if not norm_pix:
        # the two high-res patch-embed tensors get no gradient on a 3D-only step (quirk Q13)
        assert set(sd) - set(grads) == {"high_res_patch_embed.proj.weight", "high_res_patch_embed.proj.bias"}


@needs_ref
def test_oracle_vs_reference_tiefree_unpatched_argsort():
    """The unmodified reference (its own torch.argsort) on tie-free noise == oracle, fwd + bwd."""
    sd, vol, _ = toy_inputs()
    noise = O.synthetic_noise(2, 64, seed=5, tie_free=True)
    m = R.build_reference(**TOY.ref_kwargs())
    m.load_state_dict(sd, strict=True)
    ref = R.run_reference(m, vol, noise, 0.9, frame_loss=True, backward=True)
    (out, grads) = O.forward_backward(TOY, sd, vol, 0.9, noise, frame_loss=True)
    (loss, fl), pred, mask = out
    assert torch.equal(mask, ref["mask"])
    assert abs(float(loss) - float(ref["loss"])) < 1e-6
    assert _rel(pred.detach(), ref["pred"].detach()) < 1e-6
    assert set(grads) == set(ref["grads"])
    for k in grads:
        assert _rel(grads[k], ref["grads"][k]) < 2e-5, k


@needs_ref
def test_oracle_vs_reference_highres_2d_branch():
    """cfg-4 shape family: [B,1,3,128,128] 'high-res' input -> T'=1 'none' temporal path, raw spatial table."""
    sd, _, _ = toy_inputs()
    vol = O.synthetic_volume(2, 3, 128, 128, seed=3, zero_pad_frames=0)
    noise = O.synthetic_noise(2, 64, seed=6, tie_free=True)
    m = R.build_reference(**TOY.ref_kwargs())
    m.load_state_dict(sd, strict=True)
    ref = R.run_reference(m, vol, noise, 0.75, backward=True)
    (out, grads) = O.forward_backward(TOY, sd, vol, 0.75, noise)
    loss, pred, mask = out
    assert torch.equal(mask, ref["mask"])
    assert abs(float(loss) - float(ref["loss"])) < 1e-6
    assert set(grads) == set(ref["grads"])  # quirk Q13, 2D-512-only step
    for k in grads:
        assert _rel(grads[k], ref["grads"][k]) < 2e-5, k


@needs_ref
def test_cpu_baseline_restatement_vs_reference_nonflash():
    """oracle.cpu_baseline_forward == reference class with use_flash_attn=False (after key surgery)."""
    sd, vol, _ = toy_inputs()
    noise = O.synthetic_noise(2, 64, seed=5, tie_free=True)
    m = R.build_reference(flash_semantics=False, **TOY.ref_kwargs())
    ref_sd = {}
    for k, v in sd.items():
        if ".mixer.Wqkv." in k:
            for i, n in enumerate("qkv"):
                ref_sd[k.replace("mixer.Wqkv", f"attn.{n}")] = v.chunk(3, 0)[i]
        elif ".mixer.out_proj." in k:
            ref_sd[k.replace("mixer.out_proj", "attn.proj")] = v
        else:
            ref_sd[k] = v
    m.load_state_dict(ref_sd, strict=True)
    with torch.no_grad():
        ref = R.run_reference(m, vol, noise, 0.9)
        loss, pred, mask = O.cpu_baseline_forward(TOY, sd, vol, 0.9, noise)
    assert torch.equal(mask, ref["mask"])
    assert abs(float(loss) - float(ref["loss"])) < 1e-6
    assert _rel(pred, ref["pred"]) < 1e-6


# ---------------------------------------------------------------- 2D twin (OCTCube/models_mae_flash_attn.py)
@pytest.mark.parametrize("norm_pix", [False, True])
def test_toy2d_step_golden(golden_dir, norm_pix):
    g = np.load(os.path.join(golden_dir, "toy2d_step.npz"))
    tag = "np::" if norm_pix else ""
    cfg = O2.MAE2DConfig(**{**TOY2D.__dict__, "norm_pix_loss": norm_pix})
    sd, imgs, noise = toy2d_inputs()
    assert np.array_equal(imgs.numpy(), g["images"]) and np.array_equal(noise.numpy(), g["noise"])
    for k, v in sd.items():
        assert np.array_equal(v.numpy(), g["w::" + k]), k
    (loss, pred, mask, frame_loss), grads = O2.forward_backward(cfg, sd, imgs, 0.75, noise)
    assert abs(float(loss) - float(g[tag + "loss"])) <= 1e-6 * abs(float(g[tag + "loss"]))
    assert np.array_equal(mask.numpy(), g[tag + "mask"]) and float(mask.sum()) == 2 * (16 - 4)
    assert _rel(frame_loss.detach(), g[tag + "frame_loss"]) < 1e-6
    if not norm_pix:
        assert _rel(pred.detach(), g["pred"]) < 1e-6
        # the two sin-cos tables are frozen (models_mae_flash_attn.py:97,143); everything else gets a gradient
        assert set(sd) - set(grads) == set(O2.FROZEN)
        assert {k[3:] for k in g.files if k.startswith("g::")} == set(grads)
    n = 0
    for k in g.files:
        if k.startswith(tag + "g::"):
            assert _rel(grads[k[len(tag) + 3:]], g[k]) < 2e-5, k
            n += 1
    assert n >= 5


def test_2d_patch_order_and_conv_view():
    imgs = torch.randn(2, 3, 64, 64)
    p = O2.patchify(imgs, 16)
    assert p.shape == (2, 16, 768) and torch.equal(O2.unpatchify(p, 16), imgs)
    # element order is (p, q, c): the first three values of a patch are the three channels of its top-left pixel
    assert torch.equal(p[0, 5, :3], imgs[0, :, 16, 16])
    # a [B,3,H,W] image is a [B,1,3,H,W] volume with t_patch 3 for the patch-embed GEMM (same K = 768 order as the
    # Conv2d weight [E,3,16,16]) — the identity the CUDA path relies on
    w, b = torch.randn(32, 3, 16, 16), torch.randn(32)
    want = torch.nn.functional.conv2d(imgs, w, b, stride=16).flatten(2).transpose(1, 2)
    got = O.patchify(imgs.view(2, 1, 3, 64, 64), 16, 3) @ w.view(32, -1).t() + b
    assert torch.allclose(got, want, atol=1e-4, rtol=1e-4)


@needs_ref
def test_oracle2d_vs_reference_tiefree_unpatched_argsort():
    """The unmodified 2D reference (its own torch.argsort, its own sin-cos tables) on tie-free noise == oracle."""
    cfg = TOY2D
    m = R.build_reference_2d(**cfg.ref_kwargs())
    fresh = O2.init_state_dict(cfg)
    assert torch.equal(m.pos_embed, fresh["pos_embed"]) and torch.equal(m.decoder_pos_embed, fresh["decoder_pos_embed"])
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == {k: tuple(v.shape) for k, v in fresh.items()}
    sd, imgs, _ = toy2d_inputs()
    noise = O.synthetic_noise(2, cfg.num_patches, seed=5, tie_free=True)
    m.load_state_dict(sd, strict=True)
    ref = R.run_reference_2d(m, imgs, noise, 0.75, backward=True)
    (loss, pred, mask, frame_loss), grads = O2.forward_backward(cfg, sd, imgs, 0.75, noise)
    assert torch.equal(mask, ref["mask"])
    assert abs(float(loss) - float(ref["loss"])) < 1e-6
    assert _rel(pred.detach(), ref["pred"].detach()) < 1e-6 and _rel(frame_loss.detach(), ref["frame_loss"].detach()) < 1e-6
    assert set(grads) == set(ref["grads"])
    for k in grads:
        assert _rel(grads[k], ref["grads"][k]) < 2e-5, k


# ---------------------------------------------------------------- encoder-only ViT (OCTCube/models_vit_st_flash_attn.py)
@pytest.mark.parametrize("kind", ["sep", "joint"])
def test_toy_vit_golden(golden_dir, kind):
    g = np.load(os.path.join(golden_dir, "toy_vit_step.npz"))
    cfg = TOY_VIT[kind]
    sd, vol, dlogits = toy_vit_inputs(kind)
    assert np.array_equal(vol.numpy(), g[kind + "::volume"]) and np.array_equal(dlogits.numpy(), g[kind + "::dlogits"])
    (logits, emb), grads = OV.forward_backward(cfg, sd, vol, dlogits)
    assert _rel(logits.detach(), g[kind + "::logits"]) < 1e-6 and _rel(emb.detach(), g[kind + "::embedding"]) < 1e-6
    with torch.no_grad():
        assert _rel(OV.forward(cfg, sd, vol, hidden_states=True)[-1], g[kind + "::hidden_last"]) < 1e-6
    want = {k[len(kind) + 5:]: g[k] for k in g.files if k.startswith(kind + "::g::")}
    assert set(want) == set(grads) == set(sd) - {"norm.weight", "norm.bias"}   # the final norm is dead code (:249)
    for k in want:
        assert _rel(grads[k], want[k]) < 2e-5, k


@needs_ref
@pytest.mark.parametrize("sep,cls,gp", [(True, True, True), (False, True, False), (True, False, True)])
def test_vit_oracle_vs_reference(sep, cls, gp):
    cfg = OV.ViTConfig(num_frames=12, t_patch_size=3, img_size=64, num_classes=5, embed_dim=64, depth=2, num_heads=2,
                       sep_pos_embed=sep, cls_embed=cls, global_pool=gp)
    m = R.build_reference_vit(**cfg.ref_kwargs())
    sd = OV.init_state_dict(cfg, seed=4)
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == {k: tuple(v.shape) for k, v in sd.items()}
    m.load_state_dict(sd, strict=True)
    vol = O.synthetic_volume(2, 12, 64, 64, seed=9, zero_pad_frames=1)
    dlogits = torch.randn(2, 5, generator=torch.Generator().manual_seed(1))
    logits, emb = m(vol, return_embeddings=True)
    logits.backward(dlogits)
    ref_g = {k: p.grad for k, p in m.named_parameters() if p.grad is not None}
    (l2, e2), g2 = OV.forward_backward(cfg, sd, vol, dlogits)
    assert _rel(l2.detach(), logits.detach()) < 1e-6 and _rel(e2.detach(), emb.detach()) < 1e-6
    assert set(g2) == set(ref_g)
    for k in g2:
        assert _rel(g2[k], ref_g[k]) < 2e-5, k


@pytest.mark.slow
def test_full_cfg1_golden(golden_dir):
    """BASELINE cfg-1 (ViT-L, 1x48x256x256, mask 0.9, fp32 CPU): oracle reproduces the reference's loss."""
    g = json.load(open(os.path.join(golden_dir, "full_cfg1.json")))
    cfg = O.MAEConfig(num_frames=48, pred_t_dim=48)
    sd = O.init_state_dict(cfg, seed=0)
    assert sum(v.numel() for v in sd.values()) == g["n_params"] == 331626240
    vol = O.synthetic_volume(1, 48, 256, 256, seed=0)
    noise = O.synthetic_noise(1, 4096, seed=1)
    torch.set_num_threads(os.cpu_count())
    with torch.no_grad():
        (loss, fl), pred, mask = O.forward(cfg, sd, vol, 0.9, noise, frame_loss=True)
    assert float(mask.sum()) == g["mask_sum"] == 4096 - 409
    assert abs(float(loss) - g["loss"]) < 2e-5 * g["loss"]
    assert _rel(fl.flatten(), g["frame_losses"]) < 2e-5
    assert _rel(pred.flatten()[:8], g["pred_first8"]) < 1e-4
