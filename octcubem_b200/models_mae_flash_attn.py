"""B200-native 2D masked autoencoder with the module / operator surface of the reference's
OCTCube/models_mae_flash_attn.py (class MaskedAutoencoderViT :70-449, PatchEmbed :48-68, factories :451-461) — the
"2D twins" row of SURVEY §8a.

Same constructor kwargs (unknown ones are swallowed), attributes (`patch_embed.{input_size,patch_size,num_patches}`,
`input_size`, `embed_dim`, `depth`, `global_pool`), state_dict keys / shapes (`cls_token`, `pos_embed`, `mask_token`,
`decoder_pos_embed`, `patch_embed.proj.*` as a Conv2d, `blocks.i.mixer.*` ...), methods and return arities:
    patchify / unpatchify / random_masking (3-tuple) / forward_encoder / forward_decoder / forward_loss /
    forward(imgs, mask_ratio, return_frame_loss) -> (loss, pred, mask[, frame_loss]) / load_state_dict_to_backbone(_retfound)
It runs on the SAME kernels as the 3D model (octcubem_b200/models_mae.py): a [B,3,H,W] image is a [B,1,3,H,W] volume with
t_patch 3 for the im2col-free patch embedding (the Conv2d weight [E,3,16,16] flattens to the same K = 768 order), the
frozen sin-cos table takes the place of the spatial pos table (adding it before or after the keep-gather is the same
sum), the decoder's per-sample cls row rides on oct_unshuffle (y_row0 = 1) and the loss reads the (p, q, c)-ordered
target in place (OCT_LOSS_CHANNEL_LAST).  No torch op computes on the step path.
"""
from __future__ import annotations

import re
from collections import OrderedDict
from functools import partial

import numpy as np
import torch
import torch.nn as nn

from . import ops
from ._lib import LOSS_ALL_TOKENS, LOSS_CHANNEL_LAST
from .models_mae import Block, _Ctx, len_keep_of


def _to_2tuple(x):
    return tuple(x) if isinstance(x, (tuple, list)) else (x, x)


def get_2d_sincos_pos_embed(embed_dim, grid_size, cls_token=False):
    """OCTCube/util/pos_embed.py:20-68 -> numpy [(1+)grid*grid, embed_dim]: channels [0, E/2) encode the column index,
    [E/2, E) the row index, each as [sin | cos] of pos / 10000^(k / (E/4))."""
    assert embed_dim % 4 == 0

    def axis_table(pos):
        omega = np.arange(embed_dim // 4, dtype=np.float32)
        omega /= embed_dim / 4.0
        omega = 1.0 / 10000 ** omega
        ang = np.einsum("m,d->md", pos.reshape(-1), omega)
        return np.concatenate([np.sin(ang), np.cos(ang)], axis=1)

    rows, cols = np.meshgrid(np.arange(grid_size, dtype=np.float32), np.arange(grid_size, dtype=np.float32), indexing="ij")
    emb = np.concatenate([axis_table(cols), axis_table(rows)], axis=1)
    if cls_token:
        emb = np.concatenate([np.zeros([1, embed_dim]), emb], axis=0)
    return emb


class PatchEmbed(nn.Module):
    """Image to patch embedding (models_mae_flash_attn.py:48-68): Conv2d(kernel = stride = patch) -> [B, h*w, E].
    `proj` is a parameter container; the convolution runs as the tcgen05 patch-embed kernel (ops.PatchEmbedFn)."""

    def __init__(self, input_size=224, patch_size=16, in_chans=3, embed_dim=768):
        super().__init__()
        input_size, patch_size = _to_2tuple(input_size), _to_2tuple(patch_size)
        self.input_size = input_size
        self.patch_size = patch_size
        self.num_patches = (input_size[1] // patch_size[1]) * (input_size[0] // patch_size[0])
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)
        self.act_dtype = torch.bfloat16

    def forward(self, x):
        B, C, H, W = x.shape
        assert H == self.input_size[0] and W == self.input_size[1], (
            f"Input image size ({H}*{W}) doesn't match model ({self.input_size[0]}*{self.input_size[1]}).")
        return ops.PatchEmbedFn.apply(x.contiguous().float().view(B, 1, C, H, W), self.proj.weight, self.proj.bias,
                                      self.patch_size[0], C, self.act_dtype)


class MaskedAutoencoderViT(nn.Module):
    """Masked Autoencoder with VisionTransformer backbone (2D, fixed sin-cos pos-embed, cls token)."""

    def __init__(self, input_size=224, patch_size=16, in_chans=3, embed_dim=1024, depth=24, num_heads=16,
                 decoder_embed_dim=512, decoder_depth=8, decoder_num_heads=16, mlp_ratio=4.0, norm_layer=nn.LayerNorm,
                 norm_pix_loss=False, global_pool=True, cls_embed=True, use_flash_attn=True, no_qkv_bias=False,
                 qk_scale=None, drop_rate=0.0, attn_drop_rate=0.0, drop_path_rate=0.0, precision="bf16", **kwargs):
        super().__init__()
        if not use_flash_attn:
            raise NotImplementedError("octcubem_b200 implements the flash-attn variant (prenorm Block with fp32 residual, "
                                      "models_mae_flash_attn.py:100-128); the timm Block variant is out of scope")
        if drop_rate or attn_drop_rate or drop_path_rate:
            raise NotImplementedError("dropout / stochastic depth are 0 in the reference recipe and not implemented")
        if no_qkv_bias:
            raise NotImplementedError("no_qkv_bias=True is not supported")
        if in_chans != 3:
            raise NotImplementedError("the 2D model's patchify / unpatchify hard-code 3 channels (models_mae_flash_attn.py:222-239)")
        self.use_flash_attn = use_flash_attn
        self.global_pool = global_pool
        self.embed_dim = embed_dim
        self.depth = depth
        self.in_chans = in_chans
        self._rt = _Ctx()

        self.patch_embed = PatchEmbed(input_size, patch_size, in_chans, embed_dim)
        num_patches = self.patch_embed.num_patches
        self.input_size = self.patch_embed.input_size
        assert self.input_size[0] == self.input_size[1] and self.patch_embed.patch_size[0] == self.patch_embed.patch_size[1]

        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, num_patches + 1, embed_dim), requires_grad=False)  # fixed sin-cos
        self.blocks = nn.ModuleList([Block(embed_dim, num_heads, mlp_ratio, True, norm_layer, self._rt) for _ in range(depth)])
        self.norm = norm_layer(embed_dim)

        self.decoder_embed = nn.Linear(embed_dim, decoder_embed_dim, bias=True)
        self.mask_token = nn.Parameter(torch.zeros(1, 1, decoder_embed_dim))
        self.decoder_pos_embed = nn.Parameter(torch.zeros(1, num_patches + 1, decoder_embed_dim), requires_grad=False)
        self.decoder_blocks = nn.ModuleList(
            [Block(decoder_embed_dim, decoder_num_heads, mlp_ratio, True, norm_layer, self._rt) for _ in range(decoder_depth)])
        self.decoder_norm = norm_layer(decoder_embed_dim)
        self.decoder_pred = nn.Linear(decoder_embed_dim, self.patch_embed.patch_size[0] ** 2 * in_chans, bias=True)
        self.norm_pix_loss = norm_pix_loss
        self.initialize_weights()
        self.set_precision(precision)

    # ------------------------------------------------------------------ configuration
    def set_precision(self, precision: str):
        assert precision in ("bf16", "fp32")
        self._rt.precision = precision
        self.patch_embed.act_dtype = self._rt.act_dtype
        return self

    @property
    def precision(self):
        return self._rt.precision

    def shadow_of(self, p):
        return self._rt.shadows.buffer_of(p)

    def shadows_current(self):
        self._rt.shadows.mark_current()

    # ------------------------------------------------------------------ init (models_mae_flash_attn.py:184-212)
    def initialize_weights(self):
        g = int(self.patch_embed.num_patches ** 0.5)
        for table in (self.pos_embed, self.decoder_pos_embed):
            table.data.copy_(torch.from_numpy(get_2d_sincos_pos_embed(table.shape[-1], g, cls_token=True)).float().unsqueeze(0))
        w = self.patch_embed.proj.weight.data
        nn.init.xavier_uniform_(w.view([w.shape[0], -1]))
        nn.init.normal_(self.cls_token, std=0.02)
        nn.init.normal_(self.mask_token, std=0.02)
        self.apply(self._init_weights)

    @staticmethod
    def _init_weights(m):
        if isinstance(m, nn.Linear):
            nn.init.xavier_uniform_(m.weight)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    # ------------------------------------------------------------------ patchify / unpatchify (:214-240; viz + API)
    def patchify(self, imgs):
        p = self.patch_embed.patch_size[0]
        assert imgs.shape[2] == imgs.shape[3] and imgs.shape[2] % p == 0
        h = w = imgs.shape[2] // p
        x = imgs.reshape(imgs.shape[0], 3, h, p, w, p).permute(0, 2, 4, 3, 5, 1)
        return x.reshape(imgs.shape[0], h * w, p * p * 3)

    def unpatchify(self, x):
        p = self.patch_embed.patch_size[0]
        h = w = int(x.shape[1] ** 0.5)
        assert h * w == x.shape[1]
        x = x.reshape(x.shape[0], h, w, p, p, 3).permute(0, 5, 1, 3, 2, 4)
        return x.reshape(x.shape[0], 3, h * p, h * p)

    # ------------------------------------------------------------------ masking (:242-267)
    @staticmethod
    def _draw_noise(N, L, device, noise=None):
        if noise is not None:
            assert tuple(noise.shape) == (N, L), f"noise must be [{N},{L}]"
            return noise.to(device=device, dtype=torch.float32).contiguous()
        return torch.rand(N, L, device=device)  # same call shape as :251 (the harness' injection trick works unchanged)

    def random_masking(self, x, mask_ratio, noise=None):
        """x [N,L,D] -> (x_masked, mask, ids_restore); stable argsort contract of SURVEY H1."""
        N, L, D = x.shape
        noise = self._draw_noise(N, L, x.device, noise)
        mask, ids_restore, ids_keep = ops.mask_sort(noise, len_keep_of(L, mask_ratio))
        x_masked = ops.GatherTokensFn.apply(x.contiguous(), ids_keep, None, None, None)
        return x_masked.to(x.dtype), mask, ids_restore

    # ------------------------------------------------------------------ encoder (:269-297)
    def forward_encoder(self, x, mask_ratio, noise=None):
        rt, pe = self._rt, self.patch_embed
        N, C, H, W = x.shape
        assert H == pe.input_size[0] and W == pe.input_size[1], (
            f"Input image size ({H}*{W}) doesn't match model ({pe.input_size[0]}*{pe.input_size[1]}).")
        L = pe.num_patches
        noise = self._draw_noise(N, L, x.device, noise)
        mask, ids_restore, ids_keep = ops.mask_sort(noise, len_keep_of(L, mask_ratio))
        pos = self.pos_embed[0]
        cls_row = (self.cls_token + self.pos_embed[:, :1, :]).reshape(-1)
        # (x + pos)[ids_keep] == x[ids_keep] + pos[ids_keep]: the fused front end adds the table after the gather
        x = ops.EmbedTokensFn.apply(x.contiguous().float().view(N, 1, C, H, W), pe.proj.weight, pe.proj.bias, ids_keep,
                                    pos[1:].contiguous(), None, cls_row, pe.patch_size[0], C, rt.act_dtype, rt.lp(pe.proj.weight))
        residual = None
        for blk in self.blocks:
            x, residual = blk(x, residual)
        # the flash branch normalises the last MLP output only; `residual` is dropped (:284-295, quirk Q1)
        x, _ = ops.AddLNFn.apply(x, None, self.norm.weight, self.norm.bias, self.norm.eps, rt.act_dtype, False)
        return x, mask, ids_restore

    # ------------------------------------------------------------------ decoder (:299-329)
    def _decoder_tokens(self, x, ids_restore):
        rt = self._rt
        x = ops.LinearFn.apply(x.contiguous(), self.decoder_embed.weight, self.decoder_embed.bias,
                               rt.lp(self.decoder_embed.weight))
        dpos = self.decoder_pos_embed[0]
        x = ops.UnshuffleFn.apply(x.contiguous(), ids_restore, self.mask_token.reshape(-1), dpos[1:].contiguous(), None,
                                  dpos[0].contiguous(), 1)
        residual = None
        for blk in self.decoder_blocks:
            x, residual = blk(x, residual)
        x, _ = ops.AddLNFn.apply(x, None, self.decoder_norm.weight, self.decoder_norm.bias, self.decoder_norm.eps,
                                 rt.act_dtype, False)
        return ops.LinearFn.apply(x, self.decoder_pred.weight, self.decoder_pred.bias, rt.lp(self.decoder_pred.weight))

    def forward_decoder(self, x, ids_restore):
        return self._decoder_tokens(x, ids_restore)[:, 1:, :]

    # ------------------------------------------------------------------ loss (:331-350)
    def _loss(self, imgs, pred_full, row0, mask, return_frame_loss):
        N, C, H, W = imgs.shape
        flags = LOSS_CHANNEL_LAST | (LOSS_ALL_TOKENS if return_frame_loss else 0)
        loss, _, loss_tok = ops.MaskedMSELossFn.apply(imgs.contiguous().float().view(N, 1, C, H, W), pred_full,
                                                      mask.contiguous(), self.patch_embed.patch_size[0], C, row0,
                                                      bool(self.norm_pix_loss), None, flags)
        if return_frame_loss:
            return loss, loss_tok.mean(dim=-1)  # per-sample mean over ALL patches (:343-345); logging only, no gradient
        return loss

    def forward_loss(self, imgs, pred, mask, return_frame_loss=False):
        return self._loss(imgs, pred.contiguous(), 0, mask, return_frame_loss)

    # ------------------------------------------------------------------ forward (:352-359)
    def forward(self, imgs, mask_ratio=0.75, return_frame_loss=False, noise=None):
        """-> (loss, pred [N, L, p*p*3], mask [N, L]) or, with return_frame_loss, (loss, pred, mask, frame_loss [N]).
        `noise` (optional [N, L]) replaces the torch.rand draw of :251 for reproducible masks."""
        self._rt.shadows.begin_step()
        latent, mask, ids_restore = self.forward_encoder(imgs, mask_ratio, noise=noise)
        pred_full = self._decoder_tokens(latent, ids_restore)
        loss = self._loss(imgs, pred_full, 1, mask, return_frame_loss)
        if return_frame_loss:
            return loss[0], pred_full[:, 1:, :], mask, loss[1]
        return loss, pred_full[:, 1:, :], mask

    # ------------------------------------------------------------------ checkpoint key surgery (:361-449)
    def _to_mixer_keys(self, sd):
        """attn.proj -> mixer.out_proj.  The reference also flattens a 4-D `patch_embed.proj.weight` "from Conv2d to Linear"
        (:362-369) although its own PatchEmbed holds a Conv2d, so that load raises a size mismatch unless the key is
        filtered out; here a patch-embed weight of either layout is reshaped to the module's Conv2d shape instead."""
        w = sd.get("patch_embed.proj.weight")
        if w is not None and w.numel() == self.patch_embed.proj.weight.numel():
            sd["patch_embed.proj.weight"] = w.reshape(self.patch_embed.proj.weight.shape)
        return OrderedDict((re.sub(r"blocks\.(\d+)\.attn\.proj\.", r"blocks.\1.mixer.out_proj.", k), v) for k, v in sd.items())

    def load_state_dict_to_backbone(self, state_dict, strict=False, filter_keys=[]):
        """timm / mae_st-style checkpoint (separate attn.q / attn.k / attn.v) -> fused mixer.Wqkv."""
        sd = self._to_mixer_keys(dict(state_dict))
        for prefix, n in (("blocks", len(self.blocks)), ("decoder_blocks", len(self.decoder_blocks))):
            for i in range(n):
                for kind in ("weight", "bias"):
                    parts = [sd.pop(f"{prefix}.{i}.attn.{name}.{kind}") for name in ("q", "k", "v")]
                    sd[f"{prefix}.{i}.mixer.Wqkv.{kind}"] = torch.cat(parts, dim=0)
        sd = {k: v for k, v in sd.items() if not any(f in k for f in filter_keys)}
        return super().load_state_dict(sd, strict=strict)

    def load_state_dict_to_backbone_retfound(self, state_dict, strict=False, filter_keys=[], encoder_only=False):
        """RETFound / MAE checkpoint (fused attn.qkv) -> mixer.Wqkv (:403-449)."""
        sd = self._to_mixer_keys(dict(state_dict))
        groups = [("blocks", len(self.blocks))] + ([] if encoder_only else [("decoder_blocks", len(self.decoder_blocks))])
        for prefix, n in groups:
            for i in range(n):
                for kind in ("weight", "bias"):
                    sd[f"{prefix}.{i}.mixer.Wqkv.{kind}"] = sd.pop(f"{prefix}.{i}.attn.qkv.{kind}")
        sd = {k: v for k, v in sd.items() if not any(f in k for f in filter_keys)}
        return super().load_state_dict(sd, strict=strict)


# ---------------------------------------------------------------------------------------------------------------
# factories (:451-461)
# ---------------------------------------------------------------------------------------------------------------
def mae_vit_large_patch16_dec512d8b(**kwargs):
    return MaskedAutoencoderViT(patch_size=16, embed_dim=1024, depth=24, num_heads=16, decoder_embed_dim=512,
                                decoder_depth=8, decoder_num_heads=16, mlp_ratio=4,
                                norm_layer=partial(nn.LayerNorm, eps=1e-6), **kwargs)


mae_vit_large_patch16 = mae_vit_large_patch16_dec512d8b  # decoder: 512 dim, 8 blocks
