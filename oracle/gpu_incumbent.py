"""GPU incumbent = GPU bf16 oracle of the 3D-MAE step.  TEST / MEASUREMENT INFRASTRUCTURE, NOT PRODUCT CODE.

What the reference runs on a GPU (SURVEY §2.2 "bar", §8c "GPU bf16 oracle"): MaskedAutoencoderViT's flash variant
    Pre-training/models_mae_joint_res_flash_attn.py:129-152,200-220   blocks built by flash_attn.models.vit.create_block
                                                                      (prenorm Block, residual_in_fp32, MHA with
                                                                      flash_attn_qkvpacked_func, Mlp with nn.GELU)
    Pre-training/custom_util/video_vit.py:69-83                       nn.Conv3d patch embedding + permute
    Pre-training/engine_pretrain.py:110                               the whole forward under torch.autocast
restated here on top of the `flash_attn` 2.8.3 package of the image (its Block / MHA / Mlp python modules and its compiled
FA2 kernels), cuBLASLt behind nn.Linear, cuDNN behind nn.Conv3d and ATen elementwise kernels — i.e. the stock stack, none
of this repository's kernels.  /root/reference itself cannot travel to the GPU box; the control flow around the blocks is
oracle/mae3d_oracle.py's (pinned against the unmodified reference on CPU by tests/test_oracle.py) with `block_fn` swapped
for the real flash_attn Block modules.

Uses: (i) `-m gpu` parity tests — the bf16 comparator that tells how much of a bf16 deviation from the fp32 oracle is the
reference's own bf16 noise; (ii) bench.py's "incumbent" block and tools/bench_incumbent.py — the same-box, same-config,
same-precision baseline.  Only tests/, tools/ and bench.py import this file.
"""
from __future__ import annotations

import math
import sys
import types
from functools import partial

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import mae3d_oracle as O


def _stub_timm():
    """flash_attn/models/vit.py:13 imports timm.models.helpers.named_apply (used only by its own VisionTransformer)."""
    if "timm" in sys.modules:
        return
    try:
        import timm  # noqa: F401
        return
    except Exception:
        pass
    for name in ("timm", "timm.models", "timm.models.helpers"):
        m = types.ModuleType(name)
        m.__path__ = []
        sys.modules[name] = m
    sys.modules["timm.models.helpers"].named_apply = lambda *a, **k: None


def create_block(dim, heads, mlp_ratio, eps, idx, depth):
    """The create_block(...) call of models...:131-149 (all dropouts / drop-paths 0, fused_* False)."""
    _stub_timm()
    from flash_attn.models.vit import create_block as fa_create_block
    return fa_create_block(dim, heads, mlp_ratio, True, 0.0, 0.0, drop_path1=0.0, drop_path2=0.0,
                           norm_layer=partial(nn.LayerNorm, eps=eps), act_layer=nn.GELU, use_flash_attn=True,
                           fused_bias_fc=False, fused_mlp=False, fused_dropout_add_ln=False, layer_idx=idx, n_layer=depth,
                           last_layer_subset=False)


class IncumbentMAE(nn.Module):
    """Parameter names / shapes of the reference so that one state_dict loads into the reference, the CPU oracle,
    the product module and this object.  attn: "flash_attn" (the reference's kernel) or "sdpa" (torch's fused attention —
    only for boxes where the flash_attn wheel has no kernel image for the device)."""

    def __init__(self, cfg: O.MAEConfig, attn="flash_attn"):
        super().__init__()
        self.cfg = cfg
        E, D, p, u = cfg.embed_dim, cfg.decoder_embed_dim, cfg.patch_size, cfg.t_patch_size

        def pe():
            m = nn.Module()
            m.proj = nn.Conv3d(cfg.in_chans, E, kernel_size=(u, p, p), stride=(u, p, p))
            return m
        self.patch_embed, self.high_res_patch_embed = pe(), pe()
        z = lambda *s: nn.Parameter(torch.zeros(*s))  # noqa: E731
        self.cls_token, self.decoder_cls_token = z(1, 1, E), z(1, 1, D)
        self.pos_embed_spatial, self.pos_embed_temporal = z(1, cfg.hr_grid ** 2, E), z(1, cfg.t_grid, E)
        self.pos_embed_class, self.mask_token = z(1, 1, E), z(1, 1, D)
        self.decoder_pos_embed_spatial, self.decoder_pos_embed_temporal = z(1, cfg.hr_grid ** 2, D), z(1, cfg.t_grid, D)
        self.decoder_pos_embed_class = z(1, 1, D)
        self.blocks = nn.ModuleList([create_block(E, cfg.num_heads, cfg.mlp_ratio, cfg.ln_eps, i, cfg.depth)
                                     for i in range(cfg.depth)])
        self.norm = nn.LayerNorm(E, eps=cfg.ln_eps)
        self.decoder_embed = nn.Linear(E, D)
        self.decoder_blocks = nn.ModuleList([create_block(D, cfg.decoder_num_heads, cfg.mlp_ratio, cfg.ln_eps, i, cfg.decoder_depth)
                                             for i in range(cfg.decoder_depth)])
        self.decoder_norm = nn.LayerNorm(D, eps=cfg.ln_eps)
        self.decoder_pred = nn.Linear(D, cfg.patch_dim)
        if attn == "sdpa":
            for blk in list(self.blocks) + list(self.decoder_blocks):
                blk.mixer.inner_attn = _SdpaSelfAttention()
                blk.mixer.use_flash_attn = False   # (MHA.forward then calls inner_attn(qkv) without varlen kwargs)
        self.attn = attn

    def _sd(self):
        return dict(self.named_parameters())

    def forward(self, imgs, mask_ratio, noise=None, frame_loss=False):
        """models...:669-680 through oracle.forward's control flow, the blocks being the flash_attn modules."""
        sd = self._sd()
        if noise is None:
            L = (imgs.shape[2] // self.cfg.t_patch_size) * (imgs.shape[-1] // self.cfg.patch_size) ** 2
            noise = torch.rand(imgs.shape[0], L, device=imgs.device)   # models...:350
        blocks = {"blocks": self.blocks, "decoder_blocks": self.decoder_blocks}

        def block_fn(prefix, h, residual):
            group, i = prefix.rsplit(".", 1)
            return blocks[group][int(i)](h, residual)
        return O.forward(self.cfg, sd, imgs, mask_ratio, noise, frame_loss, block_fn=block_fn)


class _SdpaSelfAttention(nn.Module):
    """Drop-in for flash_attn.modules.mha.FlashSelfAttention.forward(qkv) on torch's fused attention."""

    def forward(self, qkv, causal=None, key_padding_mask=None):
        q, k, v = qkv.unbind(2)                                        # [B,S,H,d]
        o = F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2))
        return o.transpose(1, 2)


def build(cfg, sd, device, attn="flash_attn"):
    m = IncumbentMAE(cfg, attn).to(device)
    m.load_state_dict(sd, strict=True)
    return m


def step(model, vol, mask_ratio, noise=None, frame_loss=False, dtype=torch.bfloat16):
    """One reference-style training step (engine_pretrain.py:110-127 + backward): autocast forward, fp32 loss, backward."""
    with torch.autocast("cuda", dtype=dtype):
        out = model(vol, mask_ratio, noise, frame_loss)
    loss = out[0][0] if frame_loss else out[0]
    loss.backward()
    return out


def forward_backward(cfg, sd, vol, mask_ratio, noise, device="cuda:0", attn="flash_attn", frame_loss=False):
    """-> ((loss, pred, mask), {name: grad}) like oracle.forward_backward, computed by the incumbent in bf16 on the GPU."""
    m = build(cfg, sd, device, attn)
    out = step(m, vol.to(device), mask_ratio, noise.to(device), frame_loss)
    torch.cuda.synchronize()
    grads = {k: p.grad.detach().float().cpu() for k, p in m.named_parameters() if p.grad is not None}
    return out, grads


def flash_attn_available(device="cuda:0"):
    """(ok, reason): the FA2 extension imports AND has a kernel image for this device (head dims 32 and 64)."""
    try:
        from flash_attn import flash_attn_qkvpacked_func
        for d in (32, 64):
            qkv = torch.randn(1, 128, 3, 2, d, device=device, dtype=torch.bfloat16, requires_grad=True)
            o = flash_attn_qkvpacked_func(qkv, 0.0, softmax_scale=d ** -0.5, causal=False)
            o.sum().backward()
        torch.cuda.synchronize()
        return True, ""
    except Exception as e:  # noqa: BLE001
        return False, f"{type(e).__name__}: {e}"[:200]


# ----------------------------------------------------------------------------------------------------------------
# measurement (bench.py "incumbent" block, tools/bench_incumbent.py)
# ----------------------------------------------------------------------------------------------------------------
def _time(fn, n, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n


def attention_kernels(B, S, H, d, device, n=10):
    """flash_attn_qkvpacked_func forward and backward alone at one shape -> (fwd ms, bwd ms)."""
    from flash_attn import flash_attn_qkvpacked_func
    qkv = (torch.randn(B, S, 3, H, d, device=device) * 0.5).bfloat16().requires_grad_(True)
    dout = torch.randn(B, S, H, d, device=device).bfloat16()
    fwd = _time(lambda: flash_attn_qkvpacked_func(qkv, 0.0, softmax_scale=d ** -0.5, causal=False), n)

    def fb():
        qkv.grad = None
        flash_attn_qkvpacked_func(qkv, 0.0, softmax_scale=d ** -0.5, causal=False).backward(dout)
    both = _time(fb, n)
    return fwd, both - fwd


def bench_block(frames, img, batch, mask_ratio, device, timed, steps):
    """The reference's GPU stack on this box, BASELINE config of the bench line: ViT-L / 512x8x16 decoder, `batch` volumes,
    bf16 autocast, forward + backward.  Eager (how the reference's loop runs) and replayed from a CUDA graph."""
    ok, why = flash_attn_available(device)
    attn = "flash_attn" if ok else "sdpa"
    cfg = O.MAEConfig(num_frames=frames, pred_t_dim=frames)
    torch.manual_seed(1234)
    m = IncumbentMAE(cfg, attn).to(device)
    for p in m.parameters():                                            # any non-degenerate weights do for timing
        if p.dim() > 1:
            nn.init.trunc_normal_(p, std=0.02)
    g = torch.Generator().manual_seed(100)
    vol = torch.rand(batch, 1, frames, img, img, generator=g)
    vol[:, :, :3] = 0; vol[:, :, -3:] = 0
    vol = vol.to(device)

    def one():
        m.zero_grad(set_to_none=True)
        step(m, vol, mask_ratio)

    for _ in range(2):
        one()
    eager_ms = timed(one, steps) / steps
    graph_ms, graph_err = None, None
    try:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            one()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            one()
        for _ in range(2):
            graph.replay()
        graph_ms = timed(graph.replay, steps) / steps
        del graph
    except Exception as e:  # noqa: BLE001
        graph_err = f"{type(e).__name__}: {e}"[:200]
        try:
            torch.cuda.synchronize()
        except Exception:  # noqa: BLE001
            pass
    out = {"what": "the reference's GPU stack on this box: flash_attn 2.8.3 Blocks (FA2 kernels) + cuBLASLt nn.Linear + cuDNN "
                   "nn.Conv3d + ATen elementwise under torch.autocast(bf16), same volumes / batch / mask ratio, forward + backward "
                   "(oracle/gpu_incumbent.py; none of this repository's kernels)",
           "attention": attn + ("" if ok else f" (flash_attn unusable here: {why})"),
           "eager_ms_per_step": eager_ms, "eager_volumes_per_s": batch / (eager_ms / 1e3),
           "graph_ms_per_step": graph_ms, "graph_volumes_per_s": None if graph_ms is None else batch / (graph_ms / 1e3)}
    if graph_err:
        out["graph_capture_failed"] = graph_err
    del m
    torch.cuda.empty_cache()
    if ok:
        try:
            Sd = (frames // 3) * (img // 16) ** 2 + 1
            Se = int((Sd - 1) * (1 - mask_ratio)) + 1
            f32, b32 = attention_kernels(batch, Sd, 16, 32, device)
            f64, b64 = attention_kernels(batch, Se, 16, 64, device)
            out["fa2_dec_attn"] = {"shape": [batch, Sd, 16, 32], "fwd_ms": f32, "bwd_ms": b32}
            out["fa2_enc_attn"] = {"shape": [batch, Se, 16, 64], "fwd_ms": f64, "bwd_ms": b64}
        except Exception as e:  # noqa: BLE001
            out["fa2_kernels"] = f"{type(e).__name__}: {e}"[:200]
    return out
