"""CPU oracle for the contrastive step's loss.  TEST INFRASTRUCTURE, NOT PRODUCT CODE — and, this round, ahead of
the product: the CUDA path for this row does not exist yet; the oracle pins what it will have to reproduce.

Restates `ClipLoss.forward` + `gather_features` of retinal-COEM/src/open_clip/loss.py:21-63,148-229 in the configuration the
reference recipe runs (`--local-loss --gather-with-grad`, src/scripts/retclip_train/train_IR_512-MAE3D-nodrop-vit-large.sh;
labels = arange, no horovod, correct_label = 0):
    all_x   = cat(all_gather(x))                      autograd-aware: its backward is a SUM reduce-scatter of the gradients
    logits_per_image  = scale * image  @ all_enface^T   [B, B*W]
    logits_per_enface = scale * enface @ all_image^T    [B, B*W]
    labels = arange(B) + B * rank
    loss_r = (CE(logits_per_image, labels) + CE(logits_per_enface, labels)) / 2
and gives the gradients in closed form, which is what a fused all-gather + logits + cross-entropy kernel has to emit:
    P  = softmax(logits_per_image),  Q = softmax(logits_per_enface),  Y = one-hot(labels)          (all [B, B*W], per rank)
    d image_r  = scale/(2B) * [ (P_r - Y_r) all_enface  +  sum_s ((Q_s - Y_s)^T enface_s)[rows of rank r] ]
    d enface_r = scale/(2B) * [ (Q_r - Y_r) all_image   +  sum_s ((P_s - Y_s)^T image_s )[rows of rank r] ]
    d scale_r  = 1/(2B) * [ <P_r - Y_r, image_r all_enface^T> + <Q_r - Y_r, enface_r all_image^T> ]
(the sums over s are the reduce-scatter; every rank back-propagates its own loss_r, DDP then averages parameter gradients).
Pinned against the UNMODIFIED reference module under a 2-rank gloo group in tests/test_clip_oracle.py.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def clip_loss_local(image, enface, all_image, all_enface, logit_scale, rank):
    """loss_r of the module docstring from already gathered features (autograd flows through whatever the inputs carry)."""
    B = image.shape[0]
    labels = torch.arange(B) + B * rank
    lpi = logit_scale * image @ all_enface.T
    lpe = logit_scale * enface @ all_image.T
    return (F.cross_entropy(lpi, labels) + F.cross_entropy(lpe, labels)) / 2


def clip_loss_and_grads(images, enfaces, logit_scale):
    """images / enfaces: lists (one entry per rank) of [B, D] features.  Returns per-rank (loss_r, d image_r, d enface_r,
    d scale_r) from the closed form above — no autograd, no process group."""
    W, B = len(images), images[0].shape[0]
    all_image, all_enface = torch.cat(images), torch.cat(enfaces)
    dP, dQ, losses = [], [], []
    for r in range(W):
        labels = torch.arange(B) + B * r
        lpi = logit_scale * images[r] @ all_enface.T
        lpe = logit_scale * enfaces[r] @ all_image.T
        losses.append((F.cross_entropy(lpi, labels) + F.cross_entropy(lpe, labels)) / 2)
        Y = F.one_hot(labels, B * W).to(lpi.dtype)
        dP.append((torch.softmax(lpi, -1) - Y) / (2 * B))     # d loss_r / d logits_per_image
        dQ.append((torch.softmax(lpe, -1) - Y) / (2 * B))
    out = []
    for r in range(W):
        rows = slice(r * B, (r + 1) * B)
        d_img = logit_scale * (dP[r] @ all_enface + sum(dQ[s].T @ enfaces[s] for s in range(W))[rows])
        d_enf = logit_scale * (dQ[r] @ all_image + sum(dP[s].T @ images[s] for s in range(W))[rows])
        d_scale = (dP[r] * (images[r] @ all_enface.T)).sum() + (dQ[r] * (enfaces[r] @ all_image.T)).sum()
        out.append((losses[r], d_img, d_enf, d_scale))
    return out
