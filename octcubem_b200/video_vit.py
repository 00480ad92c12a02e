"""PatchEmbed with the reference's module surface (Pre-training/custom_util/video_vit.py:22-83).

`proj` is an nn.Conv3d kept purely as the parameter container (state_dict keys `*.proj.weight [E,C,u,p,p]`,
`*.proj.bias [E]` and its default init); the arithmetic is the im2col-free tcgen05 GEMM of csrc/patch_embed_tc.cu
(bf16 mode) or patchify + fp32 GEMM (fp32 parity mode).  No cuDNN convolution is ever launched.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops


def _pair(x):
    return tuple(x) if isinstance(x, (tuple, list)) else (x, x)


class PatchEmbed(nn.Module):
    """Volume to patch embedding.  forward: [B, C, T, H, W] -> [B, T', h*w, E]  (video_vit.py:74-83)."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768, frames=32, t_patch_size=4):
        super().__init__()
        img_size = _pair(img_size)
        patch_size = _pair(patch_size)
        assert img_size[1] % patch_size[1] == 0
        assert img_size[0] % patch_size[0] == 0
        assert frames % t_patch_size == 0
        self.img_size = img_size
        self.patch_size = patch_size
        self.frames = frames
        self.t_patch_size = t_patch_size
        self.grid_size = img_size[0] // patch_size[0]
        self.t_grid_size = frames // t_patch_size
        self.input_size = (self.t_grid_size, img_size[0] // patch_size[0], img_size[1] // patch_size[1])
        self.num_patches = self.input_size[0] * self.input_size[1] * self.input_size[2]
        kernel = [t_patch_size] + list(patch_size)
        self.proj = nn.Conv3d(in_chans, embed_dim, kernel_size=kernel, stride=kernel)
        self.act_dtype = torch.bfloat16  # set by the owning model (precision switch)

    def forward(self, x):
        B, C, T, H, W = x.shape
        assert H == self.img_size[0] and W == self.img_size[1], (
            f"Input image size ({H}*{W}) doesn't match model ({self.img_size[0]}*{self.img_size[1]})."
        )
        if C != 1:
            raise NotImplementedError("octcubem_b200 PatchEmbed: OCT volumes are single-channel (in_chans=1)")
        out = ops.PatchEmbedFn.apply(x.contiguous().float(), self.proj.weight, self.proj.bias, self.patch_size[0],
                                     self.t_patch_size, self.act_dtype)
        return out.view(B, T // self.t_patch_size, -1, out.shape[-1])
