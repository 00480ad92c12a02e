"""Throughput of the two modules that share the pre-training step's kernels (informational; bench.py stays the headline):
  * 2D MAE (OCTCube/models_mae_flash_attn.py surface): ViT-L/16, 224x224x3, batch 64, mask 0.75, bf16 fwd+bwd
  * encoder-only 3D ViT (OCTCube/models_vit_st_flash_attn.py surface): ViT-L/16, 60x256x256 (S = 5121), batch 2, bf16 fwd+bwd
Prints one JSON object.  usage: python tools/bench_twins.py [steps]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from octcubem_b200 import models_mae_flash_attn as M2  # noqa: E402
from octcubem_b200 import models_vit_st_flash_attn as MV  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
dev = torch.device("cuda:0")
torch.manual_seed(0)


def timed(fn, n):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n


out = {}


def section(name):
    def deco(fn):
        try:
            fn()
        except Exception as e:  # noqa: keep the other section's numbers
            out[name] = {"error": f"{type(e).__name__}: {e}"}
        torch.cuda.empty_cache()
    return deco


# ---- 2D MAE: per block 24 S dim^2 + 4 S^2 dim; Se = 49 + 1, Sd = 196 + 1 (SURVEY §8d formula)
@section("mae2d_vitl_224")
def _mae2d():
    B2 = 64
    m2 = M2.mae_vit_large_patch16(input_size=224, precision="bf16").to(dev)
    imgs = torch.rand(B2, 3, 224, 224, device=dev)


    def step2d():
        m2.zero_grad(set_to_none=True)
        loss, _, _ = m2(imgs, mask_ratio=0.75)
        loss.backward()
        return loss


    ms = timed(step2d, steps)
    Se, Sd = 50, 197
    gf = 3 * (24 * (24 * Se * 1024 ** 2 + 4 * Se ** 2 * 1024) + 8 * (24 * Sd * 512 ** 2 + 4 * Sd ** 2 * 512)
              + 2 * 196 * 768 * 1024 + 2 * Se * 1024 * 512 + 2 * Sd * 512 * 768) / 1e9
    out["mae2d_vitl_224"] = {"batch": B2, "ms_per_step": ms, "images_per_s": B2 / ms * 1e3, "tflops": gf * B2 / ms,
                             "loss": float(step2d().detach()), "eager": True}



# ---- encoder-only ViT-L at the fine-tuning length
@section("vit_st_vitl_60x256x256")
def _vit():
    BV = 2
    mv = MV.flash_attn_vit_large_patch16(num_frames=60, t_patch_size=3, img_size=256, num_classes=2, sep_pos_embed=True,
                                         cls_embed=True, global_pool=True, dropout=0.0).to(dev)
    for p in (mv.pos_embed_spatial, mv.pos_embed_temporal, mv.cls_token):
        torch.nn.init.normal_(p, std=0.02)
    vol = torch.rand(BV, 1, 60, 256, 256, device=dev)
    target = torch.tensor([0, 1], device=dev)


    def stepv():
        mv.zero_grad(set_to_none=True)
        logits = mv(vol)
        loss = torch.nn.functional.cross_entropy(logits, target)
        loss.backward()
        return loss


    ms = timed(stepv, steps)
    S = 5121
    gf = 3 * (24 * (24 * S * 1024 ** 2 + 4 * S ** 2 * 1024) + 2 * 5120 * 768 * 1024) / 1e9
    out["vit_st_vitl_60x256x256"] = {"batch": BV, "S": S, "ms_per_step": ms, "volumes_per_s": BV / ms * 1e3, "tflops": gf * BV / ms,
                                     "loss": float(stepv().detach()), "eager": True}
    with torch.no_grad():
        ms_inf = timed(lambda: mv(vol), steps)
    out["vit_st_vitl_60x256x256"]["inference_ms"] = ms_inf
    out["vit_st_vitl_60x256x256"]["inference_volumes_per_s"] = BV / ms_inf * 1e3


print(json.dumps(out))
