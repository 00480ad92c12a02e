"""CPU: host-side surface of the 2D twin module (construction is device-free; compute needs the GPU)."""
import pytest
import torch

from octcubem_b200 import models_mae_flash_attn as M2
from oracle import mae2d_oracle as O2
from oracle.gen_golden import TOY2D


def build(**kw):
    return M2.MaskedAutoencoderViT(**{**TOY2D.ref_kwargs(), **kw}, norm_layer=lambda d: torch.nn.LayerNorm(d, eps=1e-6))


def test_state_dict_surface_matches_reference_layout():
    m = build(some_unknown_argparse_flag=1)                      # unknown kwargs are swallowed like the reference's **kwargs
    want = O2.init_state_dict(TOY2D)
    got = m.state_dict()
    assert {k: tuple(v.shape) for k, v in got.items()} == {k: tuple(v.shape) for k, v in want.items()}
    assert torch.equal(got["pos_embed"], want["pos_embed"]) and torch.equal(got["decoder_pos_embed"], want["decoder_pos_embed"])
    frozen = {k for k, p in m.named_parameters() if not p.requires_grad}
    assert frozen == set(O2.FROZEN)                             # models_mae_flash_attn.py:97,143
    assert m.patch_embed.num_patches == 16 and m.patch_embed.input_size == (64, 64) and m.input_size == (64, 64)
    m.load_state_dict(want, strict=True)


def test_patchify_roundtrip_matches_oracle():
    m = build()
    imgs = torch.randn(2, 3, 64, 64)
    p = m.patchify(imgs)
    assert torch.equal(p, O2.patchify(imgs, 16)) and torch.equal(m.unpatchify(p), imgs)


def test_unsupported_variants_fail_loudly():
    for kw in ({"use_flash_attn": False}, {"drop_path_rate": 0.1}, {"no_qkv_bias": True}, {"in_chans": 1}):
        with pytest.raises(NotImplementedError):
            build(**kw)


def test_no_cpu_fallback():
    m = build()
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.rand(2, 3, 64, 64), mask_ratio=0.75)


def test_checkpoint_key_surgery():
    m = build()
    sd = O2.init_state_dict(TOY2D)
    timm_style, fused = {}, {}
    for k, v in sd.items():
        if ".mixer.Wqkv." in k:
            for i, n in enumerate("qkv"):
                timm_style[k.replace("mixer.Wqkv", f"attn.{n}")] = v.chunk(3, 0)[i]
            fused[k.replace("mixer.Wqkv", "attn.qkv")] = v
        elif ".mixer.out_proj." in k:
            timm_style[k.replace("mixer.out_proj", "attn.proj")] = v
            fused[k.replace("mixer.out_proj", "attn.proj")] = v
        else:
            timm_style[k] = v
            fused[k] = v
    for loader, ck in ((m.load_state_dict_to_backbone, timm_style), (m.load_state_dict_to_backbone_retfound, fused)):
        for p in m.parameters():
            p.data.zero_()
        res = loader(dict(ck), strict=True)
        assert not res.missing_keys and not res.unexpected_keys
        for k, v in m.state_dict().items():
            assert torch.equal(v, sd[k]), k


def test_factories():
    assert M2.mae_vit_large_patch16 is M2.mae_vit_large_patch16_dec512d8b


# ---------------------------------------------------------------- encoder-only ViT surface
def test_vit_state_dict_surface_and_key_surgery():
    from octcubem_b200 import models_vit_st_flash_attn as MV
    from oracle import vit_st_oracle as OV
    from oracle.gen_golden import TOY_VIT
    for kind, cfg in TOY_VIT.items():
        m = MV.VisionTransformer(**cfg.ref_kwargs(), use_flash_attn=True, some_unknown_flag=3,
                                 norm_layer=lambda d: torch.nn.LayerNorm(d, eps=1e-6))
        sd = OV.init_state_dict(cfg)
        assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == {k: tuple(v.shape) for k, v in sd.items()}, kind
        m.load_state_dict(sd, strict=True)
        assert m.no_weight_decay() >= {"cls_token", "pos_embed_spatial"}
        timm_style = {}
        for k, v in sd.items():
            if ".mixer.Wqkv." in k:
                for i, n in enumerate("qkv"):
                    timm_style[k.replace("mixer.Wqkv", f"attn.{n}")] = v.chunk(3, 0)[i]
            elif ".mixer.out_proj." in k:
                timm_style[k.replace("mixer.out_proj", "attn.proj")] = v
            else:
                timm_style[k] = v
        for p in m.parameters():
            p.data.zero_()
        res = m.load_state_dict_to_backbone(timm_style, strict=True)
        assert not res.missing_keys and not res.unexpected_keys
        for k, v in m.state_dict().items():
            assert torch.equal(v, sd[k]), k
    with pytest.raises(NotImplementedError):
        MV.VisionTransformer(num_frames=12, t_patch_size=3, use_flash_attn=False)
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.rand(2, 1, 12, 64, 64))
