"""CPU oracle for the 2D twin of the MAE step (OCTCube/models_mae_flash_attn.py).  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain-PyTorch (CPU, fp32, autograd) restatement of
    OCTCube/models_mae_flash_attn.py   MaskedAutoencoderViT (flash variant): PatchEmbed :48-68, patchify :214-226,
                                       random_masking :242-267 (3-tuple), forward_encoder :269-297,
                                       forward_decoder :299-329, forward_loss :331-350, forward :352-359
    OCTCube/util/pos_embed.py:20-68    fixed 2D sin-cos position table
    flash_attn 2.8.3 Block / MHA / Mlp control flow (shared with oracle/mae3d_oracle.block_forward)
Differences from the 3D model that the restatement keeps: C=3 Conv2d patch embedding, the (frozen) sin-cos pos table is
added BEFORE masking, the encoder's cls token is kept and runs through decoder_embed, the patch element order is
(p, q, c) with c fastest, `return_frame_loss` is the per-sample mean over ALL patches.

Only tests/, __graft_entry__.smoke() and bench.py's cpu legs may import it.
Parity pin: the reference ships no tests / golden vectors; this file is pinned against the UNMODIFIED reference class
executed in the build container (oracle/ref_harness.build_reference_2d, tests/test_oracle.py) and against
tests/golden/toy2d_step.npz generated from that run by oracle/gen_golden.py.
"""
from __future__ import annotations

import math
from dataclasses import asdict, dataclass
from typing import Dict

import numpy as np
import torch
import torch.nn.functional as F

from .mae3d_oracle import block_forward, len_keep_of


@dataclass
class MAE2DConfig:
    """Constructor arguments of the 2D MaskedAutoencoderViT (models_mae_flash_attn.py:74-81)."""
    input_size: int = 224
    patch_size: int = 16
    in_chans: int = 3
    embed_dim: int = 1024
    depth: int = 24
    num_heads: int = 16
    decoder_embed_dim: int = 512
    decoder_depth: int = 8
    decoder_num_heads: int = 16
    mlp_ratio: float = 4.0
    norm_pix_loss: bool = False
    ln_eps: float = 1e-6

    @property
    def grid(self):
        return self.input_size // self.patch_size

    @property
    def num_patches(self):
        return self.grid * self.grid

    @property
    def patch_dim(self):
        return self.patch_size ** 2 * self.in_chans

    def ref_kwargs(self):
        d = asdict(self)
        d.pop("ln_eps")
        return d


def sincos_2d(embed_dim: int, grid: int, cls_token: bool = True) -> torch.Tensor:
    """util/pos_embed.py:20-68 — [cls +] grid*grid rows; first half of the channels encodes the column index (the
    reference's meshgrid puts w first), second half the row index; each half = [sin | cos] over 10000^(-k/(dim/4))."""
    def one_axis(dim, pos):
        omega = np.arange(dim // 2, dtype=np.float32)
        omega /= dim / 2.0
        omega = 1.0 / 10000 ** omega
        out = np.einsum("m,d->md", pos.reshape(-1), omega)
        return np.concatenate([np.sin(out), np.cos(out)], axis=1)

    ys, xs = np.meshgrid(np.arange(grid, dtype=np.float32), np.arange(grid, dtype=np.float32), indexing="ij")
    emb = np.concatenate([one_axis(embed_dim // 2, xs), one_axis(embed_dim // 2, ys)], axis=1)
    if cls_token:
        emb = np.concatenate([np.zeros([1, embed_dim]), emb], axis=0)
    return torch.from_numpy(emb).float().unsqueeze(0)


def init_state_dict(cfg: MAE2DConfig, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Same names / shapes / distributions as models_mae_flash_attn.py:84-212."""
    g = torch.Generator().manual_seed(seed)
    E, D, p, C = cfg.embed_dim, cfg.decoder_embed_dim, cfg.patch_size, cfg.in_chans

    def xavier(out_f, in_f, *view):
        a = math.sqrt(6.0 / (in_f + out_f))
        t = (torch.rand(out_f, in_f, generator=g) * 2 - 1) * a
        return t.view(*view) if view else t

    sd = {"cls_token": torch.randn(1, 1, E, generator=g) * 0.02,
          "pos_embed": sincos_2d(E, cfg.grid),
          "mask_token": torch.randn(1, 1, D, generator=g) * 0.02,
          "decoder_pos_embed": sincos_2d(D, cfg.grid)}
    K = C * p * p
    sd["patch_embed.proj.weight"] = xavier(E, K, E, C, p, p)
    sd["patch_embed.proj.bias"] = (torch.rand(E, generator=g) * 2 - 1) / math.sqrt(K)

    def block(prefix, dim):
        hid = int(dim * cfg.mlp_ratio)
        for name, shape in (("mixer.Wqkv", (3 * dim, dim)), ("mixer.out_proj", (dim, dim)), ("mlp.fc1", (hid, dim)),
                            ("mlp.fc2", (dim, hid))):
            sd[f"{prefix}.{name}.weight"] = xavier(*shape)
            sd[f"{prefix}.{name}.bias"] = torch.zeros(shape[0])
        for n in ("norm1", "norm2"):
            sd[f"{prefix}.{n}.weight"], sd[f"{prefix}.{n}.bias"] = torch.ones(dim), torch.zeros(dim)

    for i in range(cfg.depth):
        block(f"blocks.{i}", E)
    sd["norm.weight"], sd["norm.bias"] = torch.ones(E), torch.zeros(E)
    sd["decoder_embed.weight"], sd["decoder_embed.bias"] = xavier(D, E), torch.zeros(D)
    for i in range(cfg.decoder_depth):
        block(f"decoder_blocks.{i}", D)
    sd["decoder_norm.weight"], sd["decoder_norm.bias"] = torch.ones(D), torch.zeros(D)
    sd["decoder_pred.weight"], sd["decoder_pred.bias"] = xavier(cfg.patch_dim, D), torch.zeros(cfg.patch_dim)
    return sd


FROZEN = ("pos_embed", "decoder_pos_embed")  # requires_grad=False in the reference (:97,143)


def synthetic_images(B, C, H, W, seed=0):
    return torch.rand(B, C, H, W, generator=torch.Generator().manual_seed(seed))


def patchify(imgs, p):
    """models_mae_flash_attn.py:214-226: [N,C,H,W] -> [N, h*w, p*p*C], per-patch order (p, q, c)."""
    N, C, H, W = imgs.shape
    assert H == W and H % p == 0
    h = w = H // p
    x = imgs.reshape(N, C, h, p, w, p)
    return torch.einsum("nchpwq->nhwpqc", x).reshape(N, h * w, p * p * C)


def unpatchify(x, p, C=3):
    """:228-240."""
    h = w = int(x.shape[1] ** 0.5)
    assert h * w == x.shape[1]
    x = x.reshape(x.shape[0], h, w, p, p, C)
    return torch.einsum("nhwpqc->nchpwq", x).reshape(x.shape[0], C, h * p, w * p)


def random_masking(x, mask_ratio, noise):
    """:242-267 with the stable-argsort tie contract of SURVEY H1.  -> (x_masked, mask, ids_restore) + ids_keep."""
    N, L, D = x.shape
    keep = len_keep_of(L, mask_ratio)
    assert tuple(noise.shape) == (N, L)
    ids_shuffle = torch.argsort(noise, dim=1, stable=True)
    ids_restore = torch.argsort(ids_shuffle, dim=1, stable=True)
    ids_keep = ids_shuffle[:, :keep]
    x_masked = torch.gather(x, 1, ids_keep.unsqueeze(-1).expand(-1, -1, D))
    mask = torch.ones(N, L)
    mask[:, :keep] = 0
    return x_masked, torch.gather(mask, 1, ids_restore), ids_restore, ids_keep


def forward_encoder(cfg: MAE2DConfig, sd, imgs, mask_ratio, noise):
    """:269-297."""
    assert imgs.shape[-2] == cfg.input_size and imgs.shape[-1] == cfg.input_size  # :63-66
    w = sd["patch_embed.proj.weight"]
    x = F.conv2d(imgs, w, sd["patch_embed.proj.bias"], stride=w.shape[-1]).flatten(2).transpose(1, 2)
    x = x + sd["pos_embed"][:, 1:, :]
    x, mask, ids_restore, _ = random_masking(x, mask_ratio, noise)
    cls = sd["cls_token"] + sd["pos_embed"][:, :1, :]
    x = torch.cat([cls.expand(x.shape[0], -1, -1), x], 1)
    residual = None
    for i in range(cfg.depth):
        x, residual = block_forward(sd, f"blocks.{i}", x, residual, cfg.num_heads, cfg.ln_eps)
    # the flash branch normalises the last block's MLP output only; `residual` is dropped (:284-295, quirk Q1)
    x = F.layer_norm(x, (x.shape[-1],), sd["norm.weight"], sd["norm.bias"], cfg.ln_eps)
    return x, mask, ids_restore


def forward_decoder(cfg: MAE2DConfig, sd, x, ids_restore):
    """:299-329 — the cls token stays in the sequence through decoder_embed."""
    x = F.linear(x, sd["decoder_embed.weight"], sd["decoder_embed.bias"])
    N, _, D = x.shape
    mask_tokens = sd["mask_token"].repeat(N, ids_restore.shape[1] + 1 - x.shape[1], 1)
    x_ = torch.cat([x[:, 1:, :], mask_tokens], 1)
    x_ = torch.gather(x_, 1, ids_restore.unsqueeze(-1).expand(-1, -1, D))
    x = torch.cat([x[:, :1, :], x_], 1) + sd["decoder_pos_embed"]
    residual = None
    for i in range(cfg.decoder_depth):
        x, residual = block_forward(sd, f"decoder_blocks.{i}", x, residual, cfg.decoder_num_heads, cfg.ln_eps)
    x = F.layer_norm(x, (D,), sd["decoder_norm.weight"], sd["decoder_norm.bias"], cfg.ln_eps)
    x = F.linear(x, sd["decoder_pred.weight"], sd["decoder_pred.bias"])
    return x[:, 1:, :]


def forward_loss(cfg: MAE2DConfig, imgs, pred, mask, return_frame_loss=False):
    """:331-350."""
    target = patchify(imgs, cfg.patch_size)
    if cfg.norm_pix_loss:
        mean = target.mean(dim=-1, keepdim=True)
        var = target.var(dim=-1, keepdim=True)
        target = (target - mean) / (var + 1.0e-6) ** 0.5
    loss = ((pred - target) ** 2).mean(dim=-1)
    frame_loss = loss.mean(dim=-1)
    loss = (loss * mask).sum() / mask.sum()
    return (loss, frame_loss) if return_frame_loss else loss


def forward(cfg: MAE2DConfig, sd, imgs, mask_ratio=0.75, noise=None, return_frame_loss=False):
    """:352-359 -> (loss, pred, mask[, frame_loss])."""
    latent, mask, ids_restore = forward_encoder(cfg, sd, imgs, mask_ratio, noise)
    pred = forward_decoder(cfg, sd, latent, ids_restore)
    loss = forward_loss(cfg, imgs, pred, mask, return_frame_loss)
    if return_frame_loss:
        return loss[0], pred, mask, loss[1]
    return loss, pred, mask


def forward_backward(cfg: MAE2DConfig, sd, imgs, mask_ratio, noise):
    sd = {k: v.detach().clone().requires_grad_(k not in FROZEN) for k, v in sd.items()}
    out = forward(cfg, sd, imgs, mask_ratio, noise, return_frame_loss=True)
    out[0].backward()
    return out, {k: v.grad for k, v in sd.items() if v.grad is not None}
