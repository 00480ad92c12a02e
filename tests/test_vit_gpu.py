"""GPU parity of the encoder-only 3D ViT forward / backward (OCTCube/models_vit_st_flash_attn.py surface)
against the committed reference-generated fixture tests/golden/toy_vit_step.npz and the CPU oracle."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from octcubem_b200 import models_vit_st_flash_attn as MV  # noqa: E402
from octcubem_b200 import ops  # noqa: E402
from oracle import mae3d_oracle as O  # noqa: E402
from oracle import vit_st_oracle as OV  # noqa: E402
from oracle.gen_golden import TOY_VIT, toy_vit_inputs  # noqa: E402

DEV = "cuda:0"
FP32_TOL = 1e-4   # north star: within 1e-4 relative in fp32
BF16_TOL = 2e-2   # ... and 2e-2 relative in bf16


def rel(a, b):
    a, b = a.double().cpu(), torch.as_tensor(b).double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def build(cfg, sd, precision):
    m = MV.VisionTransformer(**cfg.ref_kwargs(), use_flash_attn=True, precision=precision,
                             norm_layer=lambda d: torch.nn.LayerNorm(d, eps=cfg.ln_eps)).to(DEV).eval()
    m.load_state_dict(sd, strict=True)
    return m


@pytest.mark.parametrize("kind", ["sep", "joint"])
@pytest.mark.parametrize("precision,tol", [("fp32", FP32_TOL), ("bf16", BF16_TOL)])
def test_toy_vit_vs_reference_golden(golden_dir, kind, precision, tol):
    g = np.load(os.path.join(golden_dir, "toy_vit_step.npz"))
    sd, vol, dlogits = toy_vit_inputs(kind)
    m = build(TOY_VIT[kind], sd, precision)
    logits, emb = m(vol.to(DEV), return_embeddings=True)
    logits.backward(dlogits.to(DEV))
    assert rel(logits.float(), g[kind + "::logits"]) < tol and rel(emb.float(), g[kind + "::embedding"]) < tol
    grads = {k: p.grad for k, p in m.named_parameters()}
    assert grads["norm.weight"] is None and grads["norm.bias"] is None       # `outcome = self.norm(x)` is dead (:249)
    worst = 0.0
    for k in g.files:
        if k.startswith(kind + "::g::"):
            r = rel(grads[k[len(kind) + 5:]].float(), g[k])
            worst = max(worst, r)
            assert r < tol, (k, r)
    print(f"ViT[{kind}] worst grad rel err ({precision}): {worst:.2e}")
    with torch.no_grad():
        hs = m(vol.to(DEV), hidden_states=True)
    assert len(hs) == 2 and rel(hs[-1].float(), g[kind + "::hidden_last"]) < tol


def test_vit_no_cls_and_single_frame_vs_oracle():
    """cls_embed=False (global pool still skips row 0, :248) and T' = 1 (temporal row folded into the spatial table)."""
    for cfg in (OV.ViTConfig(num_frames=12, t_patch_size=3, img_size=64, num_classes=8, embed_dim=64, depth=2, num_heads=2,
                             sep_pos_embed=True, cls_embed=False, global_pool=True),
                OV.ViTConfig(num_frames=3, t_patch_size=3, img_size=128, num_classes=8, embed_dim=64, depth=1, num_heads=1,
                             sep_pos_embed=True, cls_embed=True, global_pool=True)):
        sd = OV.init_state_dict(cfg, seed=7)
        vol = O.synthetic_volume(2, cfg.num_frames, cfg.img_size, cfg.img_size, seed=5, zero_pad_frames=0)
        dlogits = torch.randn(2, 8, generator=torch.Generator().manual_seed(2))
        (ref_logits, ref_emb), ref_g = OV.forward_backward(cfg, sd, vol, dlogits)
        m = build(cfg, sd, "fp32")
        logits, emb = m(vol.to(DEV), return_embeddings=True)
        logits.backward(dlogits.to(DEV))
        assert rel(logits, ref_logits.detach()) < FP32_TOL and rel(emb, ref_emb.detach()) < FP32_TOL
        got = {k: p.grad for k, p in m.named_parameters() if p.grad is not None}
        assert set(got) == set(ref_g)
        for k in got:
            assert rel(got[k], ref_g[k]) < FP32_TOL, k


@pytest.mark.parametrize("xdtype,odtype", [(torch.float32, torch.float32), (torch.bfloat16, torch.float32), (torch.bfloat16, torch.bfloat16)])
@pytest.mark.parametrize("B,S,C,row0,row1", [(2, 65, 64, 1, 65), (3, 1300, 1024, 1, 1300), (2, 17, 32, 0, 1), (1, 130, 1280, 5, 69)])
def test_mean_pool_fwd_bwd(xdtype, odtype, B, S, C, row0, row1):
    g = torch.Generator().manual_seed(S)
    x = torch.randn(B, S, C, generator=g).to(xdtype)
    xr = x.float().clone().requires_grad_(True)
    want = xr[:, row0:row1].mean(dim=1)
    dout = torch.randn(B, C, generator=g).to(odtype)
    want.backward(dout.float())
    xd = x.to(DEV).requires_grad_(True)
    got = ops.MeanPoolFn.apply(xd, row0, row1, odtype)
    assert got.dtype == odtype and rel(got.float(), want.detach()) < (1e-6 if odtype == torch.float32 else 4e-3)
    got.backward(dout.to(DEV))
    assert xd.grad.dtype == xdtype
    assert rel(xd.grad.float(), xr.grad) < (1e-6 if xdtype == torch.float32 else 4e-3)
    assert float(xd.grad[:, :row0].abs().sum()) == 0.0 and float(xd.grad[:, row1:].abs().sum()) == 0.0


@pytest.mark.slow
def test_vit_large_full_length_sequence_properties():
    """ViT-L at the fine-tuning length of SURVEY §8f-3 (60 frames -> S = 5121, head_dim 64): no oracle at this size, so
    check size-independent properties — finite outputs, batch-order equivariance, gradients for everything but the dead norm."""
    m = MV.flash_attn_vit_large_patch16(num_frames=60, t_patch_size=3, img_size=256, num_classes=2, sep_pos_embed=True,
                                        cls_embed=True, global_pool=True, dropout=0.0).to(DEV).eval()
    for p in (m.pos_embed_spatial, m.pos_embed_temporal, m.cls_token):
        torch.nn.init.normal_(p, std=0.02)
    vol = O.synthetic_volume(2, 60, 256, 256, seed=1).to(DEV)
    logits, emb = m(vol, return_embeddings=True)
    assert logits.shape == (2, 2) and emb.shape == (2, 1024) and bool(torch.isfinite(logits).all())
    logits2 = m(vol.flip(0))
    assert rel(logits2.flip(0), logits.detach()) < 1e-3
    logits.square().sum().backward()
    missing = {k for k, p in m.named_parameters() if p.grad is None}
    assert missing == {"norm.weight", "norm.bias"}
    assert all(bool(torch.isfinite(p.grad).all()) for p in m.parameters() if p.grad is not None)
