"""B200-native 3D masked autoencoder with the module / operator surface of the reference's
Pre-training/models_mae_joint_res_flash_attn.py (class MaskedAutoencoderViT :29-790, factories :792-843).

Same constructor kwargs (unknown ones are swallowed, main_pretrain...:414-416 splats the whole argparse namespace),
same attributes (`patch_embed`, `high_res_patch_embed`, `input_size`, `pred_t_dim`, `t_pred_patch_size`, ...), same
state_dict keys and shapes, same methods and return arities:
    patchify / unpatchify / random_masking (4-tuple) / forward_encoder / forward_decoder / forward_loss /
    forward(imgs, mask_ratio, frame_loss, pre_mask) -> (loss, pred, mask) / forward_patch_embed /
    forward_encoder_decoder / load_state_dict_to_backbone(_retfound)
Every tensor op on the step path is a hand-written sm_100a kernel behind liboctcube_b200.so (octcubem_b200/ops.py);
nn.Linear / nn.LayerNorm / nn.Conv3d objects below are parameter containers only.

Precision: `precision="bf16"` (default) reproduces the dtype flow of the reference under
torch.autocast('cuda', torch.bfloat16) (bf16 GEMM/attention operands, fp32 accumulate, fp32 residual
stream / LayerNorm statistics / loss).  `precision="fp32"` runs fp32 CUDA-core kernels end to end (parity mode).
torch.autocast contexts around the call are ignored — the module manages dtypes itself.
"""
from __future__ import annotations

import re
from collections import OrderedDict
from functools import partial

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops, video_vit
from ._lib import OCT_BF16, OCT_F32, OCT_SIMT_BF16


def len_keep_of(L: int, mask_ratio: float) -> int:
    """models...:349 — evaluated in Python float64 and truncated (5120*(1-0.9) -> 511, 4096 -> 409)."""
    return int(L * (1 - mask_ratio))


class _Shadows:
    """bf16 copies of the GEMM weights (what autocast's weight cast produces), kept in persistent buffers so that
    CUDA-graph replays see stable addresses.  Refreshed when the fp32 master changed (or always while capturing)."""

    def __init__(self):
        self._m = {}
        self._table = None   # (key, device chunk table, chunks) of the multi-tensor cast over all registered shadows

    def get(self, p: torch.Tensor) -> torch.Tensor:
        ent = self._m.get(id(p))
        if ent is None or ent["buf"].device != p.device or ent["buf"].shape != p.shape or ent["ptr"] != p.data_ptr():
            ent = {"buf": torch.empty(p.shape, dtype=torch.bfloat16, device=p.device), "ver": None,
                   "ptr": p.data_ptr(), "fresh": False, "p": p}
            self._m[id(p)] = ent
        capturing = torch.cuda.is_current_stream_capturing()
        if ent["ver"] != p._version or (capturing and not ent["fresh"]):
            ops.cast_bf16(p.detach(), ent["buf"])  # while capturing: recorded once per forward, replayed every step
            ent["ver"] = p._version
            ent["fresh"] = True
        return ent["buf"]

    def begin_step(self):
        """Start of a forward pass.  Every shadow that get() would re-cast in this pass (master changed, or a pass that is
        being captured) is refreshed HERE by one multi-tensor launch — ~200 per-parameter launches per step otherwise —
        provided that is all of them (the usual case); anything else is left to get()."""
        ents = list(self._m.values())
        for ent in ents:
            ent["fresh"] = False
        if not ents:
            return
        capturing = torch.cuda.is_current_stream_capturing()
        if not all(ent["ptr"] == ent["p"].data_ptr() for ent in ents):
            return                           # a parameter was re-allocated: get() re-registers it
        key = tuple((ent["ptr"], ent["buf"].data_ptr()) for ent in ents)
        if self._table is None or self._table[0] != key:
            if capturing:
                return                       # the table upload is a host-to-device copy: never inside a capture
            tab, n = ops.cast_table([(ent["p"].detach(), ent["buf"]) for ent in ents])
            self._table = (key, tab, n)
        if not all(capturing or ent["ver"] != ent["p"]._version for ent in ents):
            return                           # some (or all) shadows are current: get() refreshes the others
        ops.cast_bf16_multi(self._table[1], self._table[2])
        for ent in ents:
            ent["ver"] = ent["p"]._version
            ent["fresh"] = True

    def buffer_of(self, p):
        ent = self._m.get(id(p))
        return None if ent is None or ent["ptr"] != p.data_ptr() else ent["buf"]

    def mark_current(self):
        """The shadows were just rewritten from the masters by someone else (optim.FusedAdamW): skip the next re-cast."""
        for ent in self._m.values():
            ent["ver"] = ent["p"]._version


class _Ctx:
    """Run-time switches shared by all blocks of one model."""

    def __init__(self):
        self.precision = "bf16"
        self.attn_impl = "tc"  # "tc" (tcgen05) | "simt" (CUDA-core validation kernel); bf16 mode only
        self.shadows = _Shadows()

    @property
    def act_dtype(self):
        return torch.bfloat16 if self.precision == "bf16" else torch.float32

    def lp(self, w):
        return self.shadows.get(w) if self.precision == "bf16" else None

    @property
    def attn_compute(self):
        if self.precision == "fp32":
            return OCT_F32
        return OCT_BF16 if self.attn_impl == "tc" else OCT_SIMT_BF16


class _Mixer(nn.Module):
    """Parameter layout of flash_attn.modules.mha.MHA (Wqkv, out_proj)."""

    def __init__(self, dim, num_heads, qkv_bias=True):
        super().__init__()
        assert dim % num_heads == 0
        self.num_heads = num_heads
        self.head_dim = dim // num_heads
        self.Wqkv = nn.Linear(dim, 3 * dim, bias=qkv_bias)
        self.out_proj = nn.Linear(dim, dim)


class _Mlp(nn.Module):
    """Parameter layout of flash_attn.modules.mlp.Mlp (fc1, fc2; activation = exact GELU)."""

    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.fc2 = nn.Linear(hidden, dim)


class Block(nn.Module):
    """flash_attn Block with prenorm=True, residual_in_fp32=True, all dropouts 0 (flash_attn/modules/block.py:124-192,
    built by create_block at models...:131-149):
        residual = h (+ residual);  h = norm1(residual);  h = mixer(h)
        residual = h + residual;    h = norm2(residual);  h = mlp(h);   return h, residual
    """

    def __init__(self, dim, num_heads, mlp_ratio, qkv_bias, norm_layer, rt: _Ctx):
        super().__init__()
        self.mixer = _Mixer(dim, num_heads, qkv_bias)
        self.norm1 = norm_layer(dim)
        self.mlp = _Mlp(dim, int(dim * mlp_ratio))
        self.norm2 = norm_layer(dim)
        self._rt = [rt]  # list: keep the shared context out of nn.Module's attribute registration

    def forward(self, h, residual=None):
        rt = self._rt[0]
        act = rt.act_dtype
        mx = self.mixer
        y, residual = ops.AddLNFn.apply(h, residual, self.norm1.weight, self.norm1.bias, self.norm1.eps, act, True)
        if mx.Wqkv.bias is None:
            raise NotImplementedError("no_qkv_bias=True is not supported")
        qkv = ops.LinearFn.apply(y, mx.Wqkv.weight, mx.Wqkv.bias, rt.lp(mx.Wqkv.weight))
        o = ops.AttnFn.apply(qkv, mx.num_heads, rt.attn_compute)
        h = ops.LinearFn.apply(o, mx.out_proj.weight, mx.out_proj.bias, rt.lp(mx.out_proj.weight))
        y, residual = ops.AddLNFn.apply(h, residual, self.norm2.weight, self.norm2.bias, self.norm2.eps, act, True)
        h = ops.MlpFn.apply(y, self.mlp.fc1.weight, self.mlp.fc1.bias, self.mlp.fc2.weight, self.mlp.fc2.bias,
                            rt.lp(self.mlp.fc1.weight), rt.lp(self.mlp.fc2.weight))
        return h, residual


class MaskedAutoencoderViT(nn.Module):
    """Masked Autoencoder with VisionTransformer backbone (3D, separable pos-embed, cls token)."""

    def __init__(
        self,
        input_size=256,
        patch_size=16,
        in_chans=3,
        embed_dim=1024,
        depth=24,
        num_heads=16,
        decoder_embed_dim=512,
        decoder_depth=8,
        decoder_num_heads=16,
        drop_rate=0.0,
        attn_drop_rate=0.0,
        drop_path_rate=0.0,
        mlp_ratio=4.0,
        norm_layer=nn.LayerNorm,
        norm_pix_loss=False,
        num_frames=16,
        t_patch_size=4,
        patch_embed=video_vit.PatchEmbed,
        no_qkv_bias=False,
        sep_pos_embed=False,
        trunc_init=False,
        cls_embed=False,
        pred_t_dim=8,
        high_res_input_size=512,
        use_flash_attn=False,
        precision="bf16",
        **kwargs,
    ):
        super().__init__()
        if not use_flash_attn:
            raise NotImplementedError(
                "octcubem_b200 implements the flash-attn variant (use_flash_attn=True; prenorm Block with fp32 residual, "
                "models_mae_joint_res_flash_attn.py:129-152); the timm-style video_vit.Block variant is out of scope")
        if not (sep_pos_embed and cls_embed):
            raise NotImplementedError("octcubem_b200 implements the pre-training recipe: sep_pos_embed=True, cls_embed=True")
        if drop_rate or attn_drop_rate or drop_path_rate:
            raise NotImplementedError("dropout / stochastic depth are 0 in the reference recipe and not implemented")
        if in_chans != 1:
            raise NotImplementedError("OCT volumes are single-channel: in_chans must be 1 (main_pretrain...:131)")
        self.trunc_init = trunc_init
        self.sep_pos_embed = sep_pos_embed
        self.cls_embed = cls_embed
        self.pred_t_dim = pred_t_dim
        self.in_chans = in_chans
        self.t_pred_patch_size = t_patch_size * pred_t_dim // num_frames
        self._rt = _Ctx()
        self._rt.precision = precision

        self.patch_embed = patch_embed(input_size, patch_size, in_chans, embed_dim, num_frames, t_patch_size)
        num_patches = self.patch_embed.num_patches
        self.input_size = self.patch_embed.input_size
        self.high_res_patch_embed = video_vit.PatchEmbed(high_res_input_size, patch_size, in_chans, embed_dim, num_frames,
                                                         t_patch_size)
        self.high_res_input_size = self.high_res_patch_embed.input_size
        hr_tokens = self.high_res_input_size[1] * self.high_res_input_size[2]

        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.decoder_cls_token = nn.Parameter(torch.zeros(1, 1, decoder_embed_dim))
        self.pos_embed_spatial = nn.Parameter(torch.zeros(1, hr_tokens, embed_dim))
        self.pos_embed_temporal = nn.Parameter(torch.zeros(1, self.input_size[0], embed_dim))
        self.pos_embed_class = nn.Parameter(torch.zeros(1, 1, embed_dim))

        self.use_flash_attn = use_flash_attn
        self.blocks = nn.ModuleList(
            [Block(embed_dim, num_heads, mlp_ratio, not no_qkv_bias, norm_layer, self._rt) for _ in range(depth)])
        self.norm = norm_layer(embed_dim)
        self.decoder_embed = nn.Linear(embed_dim, decoder_embed_dim, bias=True)
        self.mask_token = nn.Parameter(torch.zeros(1, 1, decoder_embed_dim))
        self.decoder_pos_embed_spatial = nn.Parameter(torch.zeros(1, hr_tokens, decoder_embed_dim))
        self.decoder_pos_embed_temporal = nn.Parameter(torch.zeros(1, self.input_size[0], decoder_embed_dim))
        self.decoder_pos_embed_class = nn.Parameter(torch.zeros(1, 1, decoder_embed_dim))
        self.decoder_blocks = nn.ModuleList(
            [Block(decoder_embed_dim, decoder_num_heads, mlp_ratio, not no_qkv_bias, norm_layer, self._rt)
             for _ in range(decoder_depth)])
        self.decoder_norm = norm_layer(decoder_embed_dim)
        self.decoder_pred = nn.Linear(decoder_embed_dim, self.t_pred_patch_size * patch_size ** 2 * in_chans, bias=True)
        self.norm_pix_loss = norm_pix_loss
        # bicubic (align_corners=False) high-res -> low-res grid resampling is a fixed linear map: build its matrix once
        hr_h, hr_w = self.high_res_input_size[1], self.high_res_input_size[2]
        eye = torch.eye(hr_h * hr_w).view(1, hr_h * hr_w, hr_h, hr_w)
        lo = F.interpolate(eye, [self.input_size[1], self.input_size[2]], mode="bicubic", align_corners=False)
        mat = lo.reshape(hr_h * hr_w, -1).t().contiguous()                 # [G_lo, G_hi], 16 non-zeros per row
        for name, t in zip(("_interp_idx", "_interp_w"), ops.ell_from_dense(mat)):
            self.register_buffer(name, t, persistent=False)
        for name, t in zip(("_interp_idx_t", "_interp_w_t"), ops.ell_from_dense(mat.t().contiguous())):
            self.register_buffer(name, t, persistent=False)
        self.initialize_weights()
        self.set_precision(precision)

    # ------------------------------------------------------------------ configuration
    def set_precision(self, precision: str):
        assert precision in ("bf16", "fp32")
        self._rt.precision = precision
        self.patch_embed.act_dtype = self._rt.act_dtype
        self.high_res_patch_embed.act_dtype = self._rt.act_dtype
        return self

    @property
    def precision(self):
        return self._rt.precision

    def set_attention_impl(self, impl: str):
        assert impl in ("tc", "simt")
        self._rt.attn_impl = impl
        return self

    # ------------------------------------------------------------------ init (models...:249-287)
    def initialize_weights(self):
        for t in (self.cls_token, self.pos_embed_spatial, self.pos_embed_temporal, self.decoder_pos_embed_spatial,
                  self.decoder_pos_embed_temporal, self.pos_embed_class, self.decoder_pos_embed_class):
            nn.init.trunc_normal_(t, std=0.02)
        w = self.patch_embed.proj.weight.data
        if self.trunc_init:
            nn.init.trunc_normal_(w)
            nn.init.trunc_normal_(self.mask_token, std=0.02)
        else:
            nn.init.xavier_uniform_(w.view([w.shape[0], -1]))
            nn.init.normal_(self.mask_token, std=0.02)
        self.apply(self._init_weights)

    def _init_weights(self, m):
        if isinstance(m, nn.Linear):
            if self.trunc_init:
                nn.init.trunc_normal_(m.weight, std=0.02)
            else:
                nn.init.xavier_uniform_(m.weight)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    # ------------------------------------------------------------------ patchify / unpatchify (viz + API; models...:289-334)
    def patchify(self, imgs, high_res=False):
        N, _, T, H, W = imgs.shape
        p = (self.high_res_patch_embed if high_res else self.patch_embed).patch_size[0]
        u = self.t_pred_patch_size
        assert W % p == 0 and H % p == 0 and T % u == 0
        h, w, t = H // p, W // p, T // u
        info = (N, T, H, W, p, u, t, h, w)
        if high_res:
            self.patch_info_high_res = info
        else:
            self.patch_info = info
        if imgs.is_cuda and imgs.shape[1] == 1 and imgs.dtype == torch.float32:
            return ops.patchify(imgs.contiguous(), p, u, torch.float32)
        x = imgs.reshape(N, self.in_chans, t, u, h, p, w, p).permute(0, 2, 4, 6, 3, 5, 7, 1)
        return x.reshape(N, t * h * w, u * p * p * self.in_chans)

    def unpatchify(self, x, high_res=False, actual_t_dim=None):
        N, T, H, W, p, u, t, h, w = self.patch_info_high_res if high_res else self.patch_info
        if actual_t_dim is not None:
            T = actual_t_dim
        x = x.reshape(N, t, h, w, u, p, p, self.in_chans).permute(0, 7, 1, 4, 2, 5, 3, 6)
        return x.reshape(N, self.in_chans, T, H, W)

    # ------------------------------------------------------------------ masking (models...:336-372)
    def _draw_noise(self, N, L, mask_ratio, device, noise=None):
        if noise is not None:
            assert tuple(noise.shape) == (N, L), f"noise must be [{N},{L}]"
            return noise.to(device=device, dtype=torch.float32).contiguous()
        if mask_ratio > 0:
            return torch.rand(N, L, device=device)  # same call shape as models...:350 (quirk Q4)
        return torch.arange(L, device=device, dtype=torch.float32).expand(N, L).contiguous()

    def random_masking(self, x, mask_ratio, pre_mask=None, noise=None):
        """x [N,L,D] -> (x_masked, mask, ids_restore, ids_keep); stable argsort contract of SURVEY H1."""
        if pre_mask is not None:
            raise NotImplementedError("pre_mask: the reference branch (models...:343-347) is dead code (quirk Q7)")
        N, L, D = x.shape
        keep = len_keep_of(L, mask_ratio)
        noise = self._draw_noise(N, L, mask_ratio, x.device, noise)
        mask, ids_restore, ids_keep = ops.mask_sort(noise, keep)
        x_masked = ops.GatherTokensFn.apply(x.contiguous(), ids_keep, None, None, None)
        return x_masked.to(x.dtype), mask, ids_restore, ids_keep

    # ------------------------------------------------------------------ pos tables
    def _spatial_table(self, table, high_res):
        """models...:416-427 / :534-544: bicubic 32x32 -> 16x16 for low-res inputs (as its fixed linear map, ops.InterpTableFn), raw for 512-px."""
        C = table.shape[-1]
        if high_res:
            return table.reshape(-1, C)
        return ops.InterpTableFn.apply(table, self._interp_idx, self._interp_w, self._interp_idx_t, self._interp_w_t)

    def _is_high_res(self, H):
        return H == self.high_res_input_size[1] * self.high_res_patch_embed.patch_size[0]

    # ------------------------------------------------------------------ encoder (models...:374-497)
    def forward_encoder(self, x, mask_ratio, pre_mask=None, noise=None):
        if pre_mask is not None:
            raise NotImplementedError("pre_mask is not forwarded by the reference's forward() (models...:677)")
        rt = self._rt
        N, C, T, H, W = x.shape
        high_res = self._is_high_res(H)
        pe = self.high_res_patch_embed if high_res else self.patch_embed
        assert H == pe.img_size[0] and W == pe.img_size[1], (
            f"Input image size ({H}*{W}) doesn't match model ({pe.img_size[0]}*{pe.img_size[1]}).")
        Tp = T // pe.t_patch_size
        G = pe.input_size[1] * pe.input_size[2]
        L = Tp * G
        keep = len_keep_of(L, mask_ratio)
        noise = self._draw_noise(N, L, mask_ratio, x.device, noise)
        mask, ids_restore, ids_keep = ops.mask_sort(noise, keep)

        pos_sp = self._spatial_table(self.pos_embed_spatial, high_res)
        pos_tmp = self.pos_embed_temporal.reshape(-1, self.pos_embed_temporal.shape[-1]) if Tp != 1 else None
        if pos_tmp is not None:
            assert pos_tmp.shape[0] == Tp, "temporal pos-embed length != T' (models...:429-436 would fail to broadcast)"
        cls_row = (self.cls_token + self.pos_embed_class).reshape(-1)
        x = ops.EmbedTokensFn.apply(x.contiguous().float(), pe.proj.weight, pe.proj.bias, ids_keep, pos_sp, pos_tmp, cls_row,
                                    pe.patch_size[0], pe.t_patch_size, rt.act_dtype, rt.lp(pe.proj.weight))
        residual = None
        for blk in self.blocks:
            x, residual = blk(x, residual)
        # quirk Q1: the final norm sees the last MLP output only; `residual` is dropped (models...:483-489)
        x, _ = ops.AddLNFn.apply(x, None, self.norm.weight, self.norm.bias, self.norm.eps, rt.act_dtype, False)
        x = x[:, 1:, :]
        return x, mask, ids_restore

    # ------------------------------------------------------------------ decoder (models...:499-606)
    def _decoder_tokens(self, x, ids_restore, high_res):
        rt = self._rt
        N = x.shape[0]
        g = self.high_res_patch_embed.grid_size if high_res else self.patch_embed.grid_size
        G = g * g
        L = ids_restore.shape[-1]
        actual_t = L // G
        x = ops.LinearFn.apply(x.contiguous(), self.decoder_embed.weight, self.decoder_embed.bias,
                               rt.lp(self.decoder_embed.weight))
        D = x.shape[-1]
        if actual_t != 1:
            pos_sp = self._spatial_table(self.decoder_pos_embed_spatial, high_res)
            pos_tmp = self.decoder_pos_embed_temporal.reshape(-1, D)
            # quirk Q10: the spatial table is repeated input_size[0] times regardless of actual_t (models...:547-548)
            assert actual_t == self.input_size[0], "decoder temporal length != input_size[0] (reference would not broadcast)"
        else:
            pos_sp = self.decoder_pos_embed_spatial.reshape(-1, D)  # raw table in the 'none' branch (models...:554-557)
            pos_tmp = None
            assert pos_sp.shape[0] == G, "T'==1 decoder path needs the high-res grid (reference would not broadcast)"
        cls_row = (self.decoder_cls_token + self.decoder_pos_embed_class).reshape(-1)
        x = ops.UnshuffleFn.apply(x.contiguous(), ids_restore, self.mask_token.reshape(-1), pos_sp.contiguous(),
                                  None if pos_tmp is None else pos_tmp.contiguous(), cls_row)
        residual = None
        for blk in self.decoder_blocks:
            x, residual = blk(x, residual)
        x, _ = ops.AddLNFn.apply(x, None, self.decoder_norm.weight, self.decoder_norm.bias, self.decoder_norm.eps,
                                 rt.act_dtype, False)
        return ops.LinearFn.apply(x, self.decoder_pred.weight, self.decoder_pred.bias, rt.lp(self.decoder_pred.weight))

    def forward_decoder(self, x, ids_restore, high_res=False):
        return self._decoder_tokens(x, ids_restore, high_res)[:, 1:, :]

    def forward_encoder_decoder(self, imgs):
        latent, mask, ids_restore = self.forward_encoder(imgs, 0)
        return self.forward_decoder(latent, ids_restore)

    # ------------------------------------------------------------------ loss (models...:613-667)
    def _loss(self, imgs, pred_full, row0, mask, frame_loss):
        T, H = imgs.shape[2], imgs.shape[-2]
        high_res = self._is_high_res(H)
        pe = self.high_res_patch_embed if high_res else self.patch_embed
        frame_idx = None
        if T != 3:
            idx = torch.linspace(0, T - 1, self.pred_t_dim).long()
            if idx.numel() != T or not torch.equal(idx, torch.arange(T)):
                frame_idx = idx.to(imgs.device)
        loss, frame_losses, _ = ops.MaskedMSELossFn.apply(imgs.contiguous().float(), pred_full, mask.contiguous(),
                                                       pe.patch_size[0], self.t_pred_patch_size, row0,
                                                       bool(self.norm_pix_loss), frame_idx)
        return (loss, frame_losses) if frame_loss else loss

    def forward_loss(self, imgs, pred, mask, frame_loss=False):
        return self._loss(imgs, pred.contiguous(), 0, mask, frame_loss)

    # ------------------------------------------------------------------ forward (models...:669-680)
    def forward(self, imgs, mask_ratio=0.75, frame_loss=False, pre_mask=None, noise=None):
        """-> (loss, pred [N, L, u*p*p], mask [N, L]); loss is (loss, frame_losses[N, T']) when frame_loss=True.
        `pre_mask` is accepted and ignored exactly like the reference (models...:677).  `noise` (optional [N, L])
        replaces the torch.rand draw of models...:350 for reproducible masks."""
        self._rt.shadows.begin_step()
        high_res = self._is_high_res(imgs.shape[-2])
        with ops.forward_pdl():
            latent, mask, ids_restore = self.forward_encoder(imgs, mask_ratio, noise=noise)
            pred_full = self._decoder_tokens(latent, ids_restore, high_res)
            loss = self._loss(imgs, pred_full, 1, mask, frame_loss)
        return loss, pred_full[:, 1:, :], mask

    # ------------------------------------------------------------------ optimizer hand-shake (optim.FusedAdamW)
    def shadow_of(self, p):
        """bf16 shadow buffer of a GEMM weight (None for parameters that have none, or before the first forward): pass
        `shadows=model.shadow_of` to optim.FusedAdamW so that the update kernel also writes the bf16 copies, then call
        `model.shadows_current()` after `optimizer.step()` — the next forward launches no weight casts."""
        return self._rt.shadows.buffer_of(p)

    def shadows_current(self):
        self._rt.shadows.mark_current()

    def forward_patch_embed(self, imgs):
        """models...:777-790 -> [N, T'*L, C]."""
        pe = self.high_res_patch_embed if self._is_high_res(imgs.shape[-2]) else self.patch_embed
        x = pe(imgs)
        N, T, L, C = x.shape
        return x.reshape(N, T * L, C)

    # ------------------------------------------------------------------ checkpoint key surgery (models...:682-775)
    @staticmethod
    def _flatten_conv2d_patch_weights(sd, keys):
        for k in keys:
            w = sd.get(k)
            if w is not None and w.dim() == 4:
                sd[k] = w.reshape(w.shape[0], -1)

    @staticmethod
    def _rename_attn_proj(sd):
        return OrderedDict((re.sub(r"blocks\.(\d+)\.attn\.proj\.", r"blocks.\1.mixer.out_proj.", k), v) for k, v in sd.items())

    def load_state_dict_to_backbone(self, state_dict, strict=False, filter_keys=[]):
        """timm/mae_st-style checkpoint (separate attn.q / attn.k / attn.v) -> fused mixer.Wqkv."""
        sd = dict(state_dict)
        self._flatten_conv2d_patch_weights(sd, ["patch_embed.proj.weight"])
        sd = self._rename_attn_proj(sd)
        for prefix, n in (("blocks", len(self.blocks)), ("decoder_blocks", len(self.decoder_blocks))):
            for i in range(n):
                for kind in ("weight", "bias"):
                    parts = [sd.pop(f"{prefix}.{i}.attn.{name}.{kind}") for name in ("q", "k", "v")]
                    sd[f"{prefix}.{i}.mixer.Wqkv.{kind}"] = torch.cat(parts, dim=0)
        sd = {k: v for k, v in sd.items() if not any(f in k for f in filter_keys)}
        return super().load_state_dict(sd, strict=strict)

    def load_state_dict_to_backbone_retfound(self, state_dict, strict=False, filter_keys=[], encoder_only=False):
        """RETFound / MAE checkpoint (fused attn.qkv) -> mixer.Wqkv."""
        sd = dict(state_dict)
        self._flatten_conv2d_patch_weights(sd, ["patch_embed.proj.weight", "high_res_patch_embed.proj.weight"])
        sd = self._rename_attn_proj(sd)
        groups = [("blocks", len(self.blocks))]
        if not encoder_only:
            groups.append(("decoder_blocks", len(self.decoder_blocks)))
        for prefix, n in groups:
            for i in range(n):
                for kind in ("weight", "bias"):
                    sd[f"{prefix}.{i}.mixer.Wqkv.{kind}"] = sd.pop(f"{prefix}.{i}.attn.qkv.{kind}")
        sd = {k: v for k, v in sd.items() if not any(f in k for f in filter_keys)}
        return super().load_state_dict(sd, strict=strict)


# ---------------------------------------------------------------------------------------------------------------
# factories (models...:792-843)
# ---------------------------------------------------------------------------------------------------------------
def flash_attn_mae_vit_large_patch16(**kwargs):
    return MaskedAutoencoderViT(patch_size=16, embed_dim=1024, depth=24, num_heads=16, mlp_ratio=4,
                                norm_layer=partial(nn.LayerNorm, eps=1e-6), use_flash_attn=True, **kwargs)


def mae_vit_base_patch16(**kwargs):
    return MaskedAutoencoderViT(patch_size=16, embed_dim=768, depth=12, num_heads=12, mlp_ratio=4,
                                norm_layer=partial(nn.LayerNorm, eps=1e-6), **kwargs)


def mae_vit_large_patch16(**kwargs):
    return MaskedAutoencoderViT(patch_size=16, embed_dim=1024, depth=24, num_heads=16, mlp_ratio=4,
                                norm_layer=partial(nn.LayerNorm, eps=1e-6), **kwargs)


def mae_vit_huge_patch14(**kwargs):
    return MaskedAutoencoderViT(patch_size=14, embed_dim=1280, depth=32, num_heads=16, mlp_ratio=4,
                                norm_layer=partial(nn.LayerNorm, eps=1e-6), **kwargs)
