"""B200-native encoder-only 3D ViT (fine-tune / inference forward) with the module surface of the reference's
OCTCube/models_vit_st_flash_attn.py (class VisionTransformer :50-297, factories :300-346) — SURVEY §8f-3.

No masking: every patch token of the [N,1,T,H,W] volume enters the encoder (S = 1 + T'·h·w, up to 5121 tokens at
E = 1024 / head_dim 64), then a global average pool without the cls token (or the cls read-out) and a Linear head.
Same constructor kwargs (unknown ones are swallowed), attributes, state_dict keys / shapes (`cls_token`,
`pos_embed_spatial` [1,h·w,E] — the model's own grid, no bicubic resampling here —, `pos_embed_temporal`,
`pos_embed_class` or the joint `pos_embed`, `patch_embed.proj.*`, `blocks.i.mixer.*`, `norm.*`, `head.*`) and
    forward(x, hidden_states=False, return_embeddings=False) -> logits | (logits, embedding) | [per-block hidden states]
It runs on the kernels of the pre-training step (octcubem_b200/models_mae.py): the fused patch-embed + pos/cls front end
with the identity keep-list, the flash-style Blocks, oct_mean_pool for the read-out.

Quirks reproduced (each verified in the reference source):
  * the flash branch hands the LAST BLOCK'S MLP OUTPUT to the pooling — the residual stream is dropped (:236-239, the same
    quirk Q1 as the MAE) — and `self.norm` is evaluated into an unused variable (`outcome`, :249), so the embedding is
    un-normalised; `norm.*` exists only as state_dict entries and receives no gradient;
  * global pooling always skips row 0 (`x[:, 1:, :]`, :248), also when cls_embed=False.
The head dropout (:165, p = 0.5 by default, active in train mode only) is torch's nn.Dropout on the [N, E] embedding: its
random stream cannot match the reference's anyway, and it is not a hot op.
"""
from __future__ import annotations

import re
from collections import OrderedDict
from functools import partial

import torch
import torch.nn as nn

from . import ops, video_vit
from .models_mae import Block, _Ctx


class VisionTransformer(nn.Module):
    """Vision Transformer with support for global average pooling (3D patches, flash-attn blocks)."""

    def __init__(self, num_frames, t_patch_size, img_size=256, patch_size=16, in_chans=1, num_classes=400, embed_dim=768,
                 depth=12, num_heads=12, mlp_ratio=4.0, no_qkv_bias=False, qk_scale=None, drop_rate=0.0, attn_drop_rate=0.0,
                 drop_path_rate=0.0, norm_layer=nn.LayerNorm, dropout=0.5, sep_pos_embed=False, cls_embed=False,
                 global_pool=False, use_flash_attn=False, precision="bf16", **kwargs):
        super().__init__()
        if not use_flash_attn:
            raise NotImplementedError("octcubem_b200 implements the flash-attn variant (use_flash_attn=True, "
                                      "models_vit_st_flash_attn.py:118-142); the video_vit.Block variant is out of scope")
        if drop_rate or attn_drop_rate or drop_path_rate:
            raise NotImplementedError("dropout / stochastic depth inside the blocks are not implemented")
        if no_qkv_bias:
            raise NotImplementedError("no_qkv_bias=True is not supported")
        if in_chans != 1:
            raise NotImplementedError("OCT volumes are single-channel: in_chans must be 1")
        self.global_pool = global_pool
        self.sep_pos_embed = sep_pos_embed
        self.cls_embed = cls_embed
        self.use_flash_attn = use_flash_attn
        self._rt = _Ctx()

        self.patch_embed = video_vit.PatchEmbed(img_size, patch_size, in_chans, embed_dim, num_frames, t_patch_size)
        num_patches = self.patch_embed.num_patches
        self.input_size = self.patch_embed.input_size
        if cls_embed:
            self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        if sep_pos_embed:
            self.pos_embed_spatial = nn.Parameter(torch.zeros(1, self.input_size[1] * self.input_size[2], embed_dim))
            self.pos_embed_temporal = nn.Parameter(torch.zeros(1, self.input_size[0], embed_dim))
            if cls_embed:
                self.pos_embed_class = nn.Parameter(torch.zeros(1, 1, embed_dim))
        else:
            self.pos_embed = nn.Parameter(torch.zeros(1, num_patches + (1 if cls_embed else 0), embed_dim), requires_grad=True)
        self.blocks = nn.ModuleList([Block(embed_dim, num_heads, mlp_ratio, True, norm_layer, self._rt) for _ in range(depth)])
        self.norm = norm_layer(embed_dim)
        self.dropout = nn.Dropout(dropout)
        self.head = nn.Linear(embed_dim, num_classes)
        nn.init.normal_(self.head.weight, std=0.02)
        self.set_precision(precision)

    def set_precision(self, precision: str):
        assert precision in ("bf16", "fp32")
        self._rt.precision = precision
        self.patch_embed.act_dtype = self._rt.act_dtype
        return self

    @property
    def precision(self):
        return self._rt.precision

    @torch.jit.ignore
    def no_weight_decay(self):
        return {"cls_token", "pos_embed", "pos_embed_spatial", "pos_embed_temporal", "pos_embed_class"}

    # ------------------------------------------------------------------ forward (:181-258)
    def forward(self, x, hidden_states=False, return_embeddings=False):
        rt, pe = self._rt, self.patch_embed
        rt.shadows.begin_step()
        N, C, T, H, W = x.shape
        assert H == pe.img_size[0] and W == pe.img_size[1], (
            f"Input image size ({H}*{W}) doesn't match model ({pe.img_size[0]}*{pe.img_size[1]}).")
        assert T == pe.frames
        Tp, G = pe.input_size[0], pe.input_size[1] * pe.input_size[2]
        L = Tp * G
        E = pe.proj.weight.shape[0]
        if self.sep_pos_embed:
            pos_sp = self.pos_embed_spatial.reshape(G, E)
            pos_tmp = self.pos_embed_temporal.reshape(Tp, E) if Tp != 1 else None
            if Tp == 1:  # the kernel's T' == 1 form has no temporal term: fold the single temporal row into the table
                pos_sp = pos_sp + self.pos_embed_temporal.reshape(1, E)
            cls_row = (self.cls_token + self.pos_embed_class).reshape(-1) if self.cls_embed else None
        else:  # one joint table (:103-113): the whole sequence is "spatial"
            table = self.pos_embed[0]
            pos_sp, pos_tmp = (table[1:] if self.cls_embed else table).contiguous(), None
            cls_row = (self.cls_token[0, 0] + table[0]) if self.cls_embed else None
        ids_all = torch.arange(L, device=x.device, dtype=torch.int64).expand(N, L).contiguous()  # nothing is masked
        x = ops.EmbedTokensFn.apply(x.contiguous().float(), pe.proj.weight, pe.proj.bias, ids_all, pos_sp.contiguous(), pos_tmp,
                                    cls_row, pe.patch_size[0], pe.t_patch_size, rt.act_dtype, rt.lp(pe.proj.weight))
        hidden, residual = [], None
        for blk in self.blocks:
            x, residual = blk(x, residual)
            hidden.append(x)
        if hidden_states:
            return hidden
        S = x.shape[1]
        # the last block's MLP output is pooled as it is: no residual, no norm (:236-251)
        row0, row1 = (1, S) if self.global_pool else (0, 1)
        embedding = ops.MeanPoolFn.apply(x.contiguous(), row0, row1, torch.float32)
        logits = ops.LinearFn.apply(self.dropout(embedding), self.head.weight, self.head.bias, None)
        if return_embeddings:
            return logits, embedding
        return logits

    # ------------------------------------------------------------------ checkpoint key surgery (:260-297)
    def load_state_dict_to_backbone(self, state_dict, strict=False, filter_keys=[]):
        """timm / mae_st-style checkpoint (separate attn.q / attn.k / attn.v, attn.proj) -> mixer.Wqkv / mixer.out_proj.
        A patch-embed weight of any layout with the right element count is reshaped to the module's Conv3d shape (the
        reference flattens 4-D weights "to Linear", :261-267, which its own Conv3d PatchEmbed cannot load)."""
        sd = dict(state_dict)
        w = sd.get("patch_embed.proj.weight")
        if w is not None and w.numel() == self.patch_embed.proj.weight.numel():
            sd["patch_embed.proj.weight"] = w.reshape(self.patch_embed.proj.weight.shape)
        sd = OrderedDict((re.sub(r"blocks\.(\d+)\.attn\.proj\.", r"blocks.\1.mixer.out_proj.", k), v) for k, v in sd.items())
        for i in range(len(self.blocks)):
            for kind in ("weight", "bias"):
                parts = [sd.pop(f"blocks.{i}.attn.{name}.{kind}") for name in ("q", "k", "v")]
                sd[f"blocks.{i}.mixer.Wqkv.{kind}"] = torch.cat(parts, dim=0)
        sd = {k: v for k, v in sd.items() if not any(f in k for f in filter_keys)}
        return super().load_state_dict(sd, strict=strict)


# ---------------------------------------------------------------------------------------------------------------
# factories (:300-346)
# ---------------------------------------------------------------------------------------------------------------
def vit_base_patch16(**kwargs):
    return VisionTransformer(patch_size=16, embed_dim=768, depth=12, num_heads=12, mlp_ratio=4,
                             norm_layer=partial(nn.LayerNorm, eps=1e-6), **kwargs)


def flash_attn_vit_large_patch16(**kwargs):
    return VisionTransformer(patch_size=16, embed_dim=1024, depth=24, num_heads=16, mlp_ratio=4,
                             norm_layer=partial(nn.LayerNorm, eps=1e-6), use_flash_attn=True, **kwargs)


def vit_large_patch16(**kwargs):
    return VisionTransformer(patch_size=16, embed_dim=1024, depth=24, num_heads=16, mlp_ratio=4,
                             norm_layer=partial(nn.LayerNorm, eps=1e-6), **kwargs)


def vit_huge_patch14(**kwargs):
    return VisionTransformer(patch_size=16, embed_dim=1280, depth=32, num_heads=16, mlp_ratio=4,
                             norm_layer=partial(nn.LayerNorm, eps=1e-6), **kwargs)
