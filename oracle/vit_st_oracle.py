"""CPU oracle for the encoder-only 3D ViT forward (OCTCube/models_vit_st_flash_attn.py).  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Plain-PyTorch (CPU, fp32, autograd) restatement of VisionTransformer.forward (:181-258) in its flash configuration
(create_block blocks :118-142, shared with oracle/mae3d_oracle.block_forward): patch-embed (util/video_vit.py PatchEmbed),
cls token, separable or joint position table, blocks, read-out, head.  Quirks kept: the last block's MLP output is pooled
without residual or norm (`outcome = self.norm(x)` is dead, :249); global pooling skips row 0 unconditionally (:248).

Only tests/ may import it.  Parity pin: pinned against the UNMODIFIED reference class executed in the build container
(oracle/ref_harness.build_reference_vit, tests/test_oracle.py) and against tests/golden/toy_vit_step.npz generated from
that run by oracle/gen_golden.py; the reference ships no tests / golden vectors of its own.
"""
from __future__ import annotations

import math
from dataclasses import asdict, dataclass
from typing import Dict

import torch
import torch.nn.functional as F

from .mae3d_oracle import block_forward, patch_embed


@dataclass
class ViTConfig:
    """Constructor arguments of VisionTransformer (models_vit_st_flash_attn.py:53-77)."""
    num_frames: int = 60
    t_patch_size: int = 3
    img_size: int = 256
    patch_size: int = 16
    in_chans: int = 1
    num_classes: int = 2
    embed_dim: int = 1024
    depth: int = 24
    num_heads: int = 16
    mlp_ratio: float = 4.0
    sep_pos_embed: bool = True
    cls_embed: bool = True
    global_pool: bool = True
    ln_eps: float = 1e-6

    @property
    def t_grid(self):
        return self.num_frames // self.t_patch_size

    @property
    def grid(self):
        return self.img_size // self.patch_size

    def ref_kwargs(self):
        d = asdict(self)
        d.pop("ln_eps")
        return d


def init_state_dict(cfg: ViTConfig, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Names / shapes of models_vit_st_flash_attn.py:84-168 (values: small random, so that every term matters)."""
    g = torch.Generator().manual_seed(seed)
    E, p, u = cfg.embed_dim, cfg.patch_size, cfg.t_patch_size

    def rn(*shape, std=0.02):
        return torch.randn(*shape, generator=g) * std

    def xavier(out_f, in_f, *view):
        a = math.sqrt(6.0 / (in_f + out_f))
        t = (torch.rand(out_f, in_f, generator=g) * 2 - 1) * a
        return t.view(*view) if view else t

    sd = {}
    if cfg.cls_embed:
        sd["cls_token"] = rn(1, 1, E)
    if cfg.sep_pos_embed:
        sd["pos_embed_spatial"] = rn(1, cfg.grid ** 2, E)
        sd["pos_embed_temporal"] = rn(1, cfg.t_grid, E)
        if cfg.cls_embed:
            sd["pos_embed_class"] = rn(1, 1, E)
    else:
        sd["pos_embed"] = rn(1, cfg.t_grid * cfg.grid ** 2 + (1 if cfg.cls_embed else 0), E)
    K = cfg.in_chans * u * p * p
    sd["patch_embed.proj.weight"] = xavier(E, K, E, cfg.in_chans, u, p, p)
    sd["patch_embed.proj.bias"] = (torch.rand(E, generator=g) * 2 - 1) / math.sqrt(K)
    hid = int(E * cfg.mlp_ratio)
    for i in range(cfg.depth):
        pre = f"blocks.{i}"
        for name, shape in (("mixer.Wqkv", (3 * E, E)), ("mixer.out_proj", (E, E)), ("mlp.fc1", (hid, E)), ("mlp.fc2", (E, hid))):
            sd[f"{pre}.{name}.weight"] = xavier(*shape)
            sd[f"{pre}.{name}.bias"] = rn(shape[0])
        for n in ("norm1", "norm2"):
            sd[f"{pre}.{n}.weight"], sd[f"{pre}.{n}.bias"] = 1 + rn(E, std=0.05), rn(E, std=0.05)
    sd["norm.weight"], sd["norm.bias"] = torch.ones(E), torch.zeros(E)
    sd["head.weight"], sd["head.bias"] = rn(cfg.num_classes, E), rn(cfg.num_classes)
    return sd


def forward(cfg: ViTConfig, sd, x, hidden_states=False):
    """:181-258 in eval mode (the head dropout is the identity) -> (logits, embedding) or the per-block hidden states."""
    assert x.shape[-2] == cfg.img_size and x.shape[-1] == cfg.img_size
    x = patch_embed(x, sd["patch_embed.proj.weight"], sd["patch_embed.proj.bias"])
    N, T, L, C = x.shape
    x = x.reshape(N, T * L, C)
    if cfg.cls_embed:
        x = torch.cat([sd["cls_token"].expand(N, -1, -1), x], 1)
    if cfg.sep_pos_embed:
        pos = sd["pos_embed_spatial"].repeat(1, cfg.t_grid, 1) + torch.repeat_interleave(sd["pos_embed_temporal"], cfg.grid ** 2, dim=1)
        if cfg.cls_embed:
            pos = torch.cat([sd["pos_embed_class"].expand(pos.shape[0], -1, -1), pos], 1)
    else:
        pos = sd["pos_embed"]
    x = x + pos
    hidden, residual = [], None
    for i in range(cfg.depth):
        x, residual = block_forward(sd, f"blocks.{i}", x, residual, cfg.num_heads, cfg.ln_eps)
        hidden.append(x)
    if hidden_states:
        return hidden
    emb = x[:, 1:, :].mean(dim=1) if cfg.global_pool else x[:, 0]  # `outcome = self.norm(x)` is never used (:249)
    return F.linear(emb, sd["head.weight"], sd["head.bias"]), emb


def forward_backward(cfg: ViTConfig, sd, x, dlogits):
    sd = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()}
    logits, emb = forward(cfg, sd, x)
    logits.backward(dlogits)
    return (logits, emb), {k: v.grad for k, v in sd.items() if v.grad is not None}
