"""octcubem_b200 — B200-native (sm_100a) implementation of the OCTCube 3D-MAE pre-training step.

Drop-in for the reference's Pre-training/models_mae_joint_res_flash_attn.py:
    from octcubem_b200 import models_mae            # models_mae.__dict__[args.model](**vars(args))
and, on the same kernels,
    from octcubem_b200 import models_mae_flash_attn      # OCTCube/models_mae_flash_attn.py (2D MAE)
    from octcubem_b200 import models_vit_st_flash_attn   # OCTCube/models_vit_st_flash_attn.py (encoder-only 3D ViT)
    from octcubem_b200 import engine_pretrain, optim, dp # the loop body of engine_pretrain.py, AdamW + schedule, DDP's role
The arithmetic lives in liboctcube_b200.so (hand-written CUDA, C ABI in include/octcube_b200.h).
"""
from . import _lib  # noqa: F401
from .models_mae import (MaskedAutoencoderViT, flash_attn_mae_vit_large_patch16, mae_vit_base_patch16,  # noqa: F401
                         mae_vit_huge_patch14, mae_vit_large_patch16)
from .video_vit import PatchEmbed  # noqa: F401

__all__ = ["MaskedAutoencoderViT", "PatchEmbed", "flash_attn_mae_vit_large_patch16", "mae_vit_base_patch16",
           "mae_vit_large_patch16", "mae_vit_huge_patch14"]
