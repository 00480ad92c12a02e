"""Generates tests/golden/* by running the UNMODIFIED reference (oracle/ref_harness.py) in the build
container.  Run:  python oracle/gen_golden.py [--full]      (test infrastructure, not product code)

Fixtures:
  toy_step.npz       toy 3D-MAE (E=64/2 heads, D=32/1 head, 12x64x64, B=2): weights, volume, noise,
                     reference loss / frame_losses / mask / pred / every parameter gradient (fp32 CPU).
  toy_step_normpix.npz   same with norm_pix_loss=True (loss + a few grads only)
  masking_cases.npz  reference random_masking on noise rows at L in {1024,4096,5120}: natural torch.rand
                     noise and 1/37-quantised noise (ties; argsort forced stable = CUDA radix-sort behaviour,
                     SURVEY H1) and tie-free noise (reference argsort untouched).
  toy2d_step.npz     the 2D twin (OCTCube/models_mae_flash_attn.py), toy size (E=64/2 heads, D=32/1 head, 3x64x64, B=2, mask
                     0.75) with norm_pix_loss off and on: images, noise, reference loss / frame_loss / mask / pred / every
                     parameter gradient.
  toy_vit_step.npz   the encoder-only ViT (OCTCube/models_vit_st_flash_attn.py), toy size (E=64/2 heads, 12x64x64, B=2, 5 classes),
                     eval mode: "sep::" = separable pos + cls + global pool, "joint::" = joint pos table + cls read-out:
                     volume, dlogits, reference logits / embedding / last hidden state / every parameter gradient.
  full_cfg1_grads.npz  ViT-L, 48x256x256, mask 0.9: reference loss + the norm of every parameter gradient + strided slices of 18
                     gradient tensors, for the first volume alone (cfg-1) and for the batch of 8 (cfg-2) (only with --full-grads).
  full_cfg3_grads.npz  the same for ONE 60x256x256 volume (BASELINE cfg-3: L = 5120, keep = 511, S_dec = 5121).
  full_cfg1.json     ViT-L, 1x48x256x256, mask 0.9 (BASELINE cfg-1): reference loss / mask sum / pred stats
                     for oracle.init_state_dict(seed 0) weights (only with --full; ~1 min).
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import mae2d_oracle as O2  # noqa: E402
from oracle import mae3d_oracle as O  # noqa: E402
from oracle import ref_harness as R  # noqa: E402
from oracle import vit_st_oracle as OV  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")

TOY = O.MAEConfig(input_size=64, patch_size=16, in_chans=1, embed_dim=64, depth=2, num_heads=2,
                  decoder_embed_dim=32, decoder_depth=1, decoder_num_heads=1, num_frames=12, t_patch_size=3,
                  pred_t_dim=12, high_res_input_size=128)


TOY2D = O2.MAE2DConfig(input_size=64, patch_size=16, in_chans=3, embed_dim=64, depth=2, num_heads=2, decoder_embed_dim=32,
                       decoder_depth=1, decoder_num_heads=1)


def toy2d_inputs():
    sd = O.perturb_state_dict(O2.init_state_dict(TOY2D, seed=0))
    return sd, O2.synthetic_images(2, 3, 64, 64, seed=0), O.synthetic_noise(2, TOY2D.num_patches, seed=1)


def gen_toy2d():
    sd, imgs, noise = toy2d_inputs()
    rec = {"images": imgs.numpy(), "noise": noise.numpy()}
    for k, v in sd.items():
        rec["w::" + k] = v.numpy()
    for norm_pix in (False, True):
        cfg = O2.MAE2DConfig(**{**TOY2D.__dict__, "norm_pix_loss": norm_pix})
        m = R.build_reference_2d(**cfg.ref_kwargs())
        m.load_state_dict(sd, strict=True)
        out = R.run_reference_2d(m, imgs, noise, 0.75, force_stable_argsort=True, backward=True)
        tag = "np::" if norm_pix else ""
        rec[tag + "loss"] = out["loss"].detach().numpy()
        rec[tag + "frame_loss"] = out["frame_loss"].detach().numpy()
        rec[tag + "mask"] = out["mask"].numpy()
        if not norm_pix:
            rec["pred"] = out["pred"].detach().numpy()
        for k, v in out["grads"].items():
            if not norm_pix or k in ("decoder_pred.weight", "blocks.0.mixer.Wqkv.weight", "cls_token", "mask_token",
                                     "patch_embed.proj.weight"):
                rec[tag + "g::" + k] = v.numpy()
        print("toy2d norm_pix", norm_pix, "loss", float(out["loss"]), "mask sum", float(out["mask"].sum()))
    np.savez_compressed(os.path.join(GOLD, "toy2d_step.npz"), **rec)


TOY_VIT = {"sep": OV.ViTConfig(num_frames=12, t_patch_size=3, img_size=64, num_classes=5, embed_dim=64, depth=2, num_heads=2,
                               sep_pos_embed=True, cls_embed=True, global_pool=True),
           "joint": OV.ViTConfig(num_frames=12, t_patch_size=3, img_size=64, num_classes=5, embed_dim=64, depth=2, num_heads=2,
                                 sep_pos_embed=False, cls_embed=True, global_pool=False)}


def toy_vit_inputs(kind):
    sd = OV.init_state_dict(TOY_VIT[kind], seed=0)
    vol = O.synthetic_volume(2, 12, 64, 64, seed=0, zero_pad_frames=1)
    dlogits = torch.randn(2, 5, generator=torch.Generator().manual_seed(3))
    return sd, vol, dlogits


def gen_toy_vit():
    rec = {}
    for kind, cfg in TOY_VIT.items():
        sd, vol, dlogits = toy_vit_inputs(kind)
        m = R.build_reference_vit(**cfg.ref_kwargs())
        m.load_state_dict(sd, strict=True)
        logits, emb = m(vol, return_embeddings=True)
        m.zero_grad(set_to_none=True)
        logits.backward(dlogits)
        with torch.no_grad():
            hidden = m(vol, hidden_states=True)
        rec[kind + "::volume"], rec[kind + "::dlogits"] = vol.numpy(), dlogits.numpy()
        rec[kind + "::logits"], rec[kind + "::embedding"] = logits.detach().numpy(), emb.detach().numpy()
        rec[kind + "::hidden_last"] = hidden[-1].numpy()
        for k, p in m.named_parameters():
            if p.grad is not None:
                rec[f"{kind}::g::{k}"] = p.grad.numpy()
        print("toy vit", kind, "logits", logits.detach().flatten()[:3].tolist(),
              "no grad:", [k for k, p in m.named_parameters() if p.grad is None])
    np.savez_compressed(os.path.join(GOLD, "toy_vit_step.npz"), **rec)


def toy_inputs():
    sd = O.perturb_state_dict(O.init_state_dict(TOY, seed=0))
    vol = O.synthetic_volume(2, 12, 64, 64, seed=0, zero_pad_frames=1)
    noise = O.synthetic_noise(2, TOY.t_grid * TOY.grid ** 2, seed=1)
    return sd, vol, noise


def gen_toy(norm_pix):
    cfg = O.MAEConfig(**{**TOY.__dict__, "norm_pix_loss": norm_pix})
    sd, vol, noise = toy_inputs()
    m = R.build_reference(**cfg.ref_kwargs())
    missing = m.load_state_dict(sd, strict=True)
    out = R.run_reference(m, vol, noise, 0.9, frame_loss=True, force_stable_argsort=True, backward=True)
    rec = {"loss": out["loss"].detach().numpy(), "frame_losses": out["frame_losses"].detach().numpy(),
           "mask": out["mask"].numpy(), "volume": vol.numpy(), "noise": noise.numpy()}
    if not norm_pix:
        rec["pred"] = out["pred"].detach().numpy()
        for k, v in sd.items():
            rec["w::" + k] = v.numpy()
        for k, v in out["grads"].items():
            rec["g::" + k] = v.numpy()
    else:
        for k in ("decoder_pred.weight", "blocks.0.mixer.Wqkv.weight", "pos_embed_spatial", "mask_token"):
            rec["g::" + k] = out["grads"][k].numpy()
    name = "toy_step_normpix.npz" if norm_pix else "toy_step.npz"
    np.savez_compressed(os.path.join(GOLD, name), **rec)
    print(name, "loss", float(out["loss"]), "mask sum", float(out["mask"].sum()))


def gen_masking():
    m = R.build_reference(**TOY.ref_kwargs())
    rec = {}
    for L in (1024, 4096, 5120):
        for kind in ("natural", "tiefree", "quantized"):
            noise = O.synthetic_noise(2, L, seed=7 + L, tie_free=(kind == "tiefree"))
            if kind == "quantized":  # many exact ties, some straddling the keep boundary
                noise = torch.floor(noise * 37.0) / 37.0
            x = torch.arange(2 * L * 2, dtype=torch.float32).view(2, L, 2)
            with R.inject_noise(noise, force_stable_argsort=(kind != "tiefree")):
                xm, mask, ids_restore, ids_keep = m.random_masking(x, 0.9)
            rec[f"{kind}_{L}_noise"] = noise.numpy()
            rec[f"{kind}_{L}_mask"] = mask.numpy().astype(np.uint8)
            rec[f"{kind}_{L}_ids_restore"] = ids_restore.numpy().astype(np.int32)
            rec[f"{kind}_{L}_ids_keep"] = ids_keep.numpy().astype(np.int32)
            print(kind, L, "keep", ids_keep.shape[1],
                  "rows with ties", int(sum(len(torch.unique(r)) < L for r in noise)))
    np.savez_compressed(os.path.join(GOLD, "masking_cases.npz"), **rec)


def gen_full():
    cfg = O.MAEConfig(num_frames=48, pred_t_dim=48)
    sd = O.init_state_dict(cfg, seed=0)
    vol = O.synthetic_volume(1, 48, 256, 256, seed=0)
    noise = O.synthetic_noise(1, cfg.t_grid * cfg.grid ** 2, seed=1)
    m = R.build_reference(**cfg.ref_kwargs())
    m.load_state_dict(sd, strict=True)
    with torch.no_grad():
        out = R.run_reference(m, vol, noise, 0.9, frame_loss=True, force_stable_argsort=True)
    rec = {"config": "ViT-L 1x48x256x256 mask 0.9 fp32 CPU, oracle.init_state_dict(seed=0), volume seed 0, noise seed 1",
           "loss": float(out["loss"]), "mask_sum": float(out["mask"].sum()),
           "frame_losses": out["frame_losses"].flatten().tolist(),
           "pred_mean": float(out["pred"].mean()), "pred_std": float(out["pred"].std()),
           "pred_first8": out["pred"].flatten()[:8].tolist(),
           "n_params": sum(v.numel() for v in sd.values())}
    json.dump(rec, open(os.path.join(GOLD, "full_cfg1.json"), "w"), indent=1)
    print(rec)


GRAD_SLICE_TENSORS = ("pos_embed_spatial", "pos_embed_temporal", "cls_token", "patch_embed.proj.weight", "blocks.0.mixer.Wqkv.weight",
                      "blocks.0.norm1.weight", "blocks.11.mlp.fc1.weight", "blocks.23.mlp.fc2.weight", "blocks.23.mixer.out_proj.bias",
                      "norm.weight", "decoder_embed.weight", "mask_token", "decoder_pos_embed_spatial",
                      "decoder_blocks.0.mixer.Wqkv.weight", "decoder_blocks.7.mlp.fc2.weight", "decoder_norm.bias",
                      "decoder_pred.weight", "decoder_pred.bias")


def grad_slice(t, n=4096):
    """Fixed, size-independent sample of a gradient tensor: every (numel // n)-th element of the flattened tensor."""
    f = t.reshape(-1)
    return f[:: max(1, f.numel() // n)][:n]


def gen_full_grads_cfg3():
    """cfg-3 volume size (60x256x256: L = 5120, keep = 511 by Python's float truncation, S_dec = 5121), one volume: reference
    loss, gradient norms and slices like gen_full_grads -> full_cfg3_grads.npz."""
    cfg = O.MAEConfig(num_frames=60, pred_t_dim=60)
    sd = O.init_state_dict(cfg, seed=0)
    vol = O.synthetic_volume(1, 60, 256, 256, seed=0)
    noise = O.synthetic_noise(1, cfg.t_grid * cfg.grid ** 2, seed=1)
    m = R.build_reference(**cfg.ref_kwargs())
    m.load_state_dict(sd, strict=True)
    out = R.run_reference(m, vol, noise, 0.9, frame_loss=True, force_stable_argsort=True, backward=True)
    rec = {"b1::loss": np.float64(float(out["loss"])), "b1::frame_losses": out["frame_losses"].detach().flatten().numpy(),
           "mask_sum_per_volume": np.float64(float(out["mask"].sum()))}
    for k, v in out["grads"].items():
        rec["b1::norm::" + k] = np.float64(v.double().norm())
        if k in GRAD_SLICE_TENSORS:
            rec["b1::slice::" + k] = grad_slice(v).float().numpy()
    np.savez_compressed(os.path.join(GOLD, "full_cfg3_grads.npz"), **rec)
    print("full_cfg3_grads.npz: loss", rec["b1::loss"], "mask sum", rec["mask_sum_per_volume"])


def gen_full_grads(batch=8):
    """Reference GRADIENTS at production size (ViT-L, 48x256x256, mask 0.9, fp32 CPU): cfg-1 (the first volume alone) and a
    batch of `batch` volumes (BASELINE cfg-2's batch).  The reference is run once per volume — its softmax(QK^T) at S = 4097
    needs ~20 GB per volume — and the batch step is assembled from the per-volume steps: every volume masks the same number of
    tokens, so the batch loss `(loss * mask).sum() / mask.sum()` (models...:664) is the mean of the per-volume losses and its
    gradient the mean of theirs.  Stored: loss, frame losses, the norm of EVERY parameter gradient and fixed strided slices
    (4096 elements) of GRAD_SLICE_TENSORS."""
    cfg = O.MAEConfig(num_frames=48, pred_t_dim=48)
    sd = O.init_state_dict(cfg, seed=0)
    vols = O.synthetic_volume(batch, 48, 256, 256, seed=0)
    noises = O.synthetic_noise(batch, cfg.t_grid * cfg.grid ** 2, seed=1)
    m = R.build_reference(**cfg.ref_kwargs())
    m.load_state_dict(sd, strict=True)
    acc, losses, frames, msum = None, [], [], None
    rec = {}
    for i in range(batch):
        out = R.run_reference(m, vols[i:i + 1], noises[i:i + 1], 0.9, frame_loss=True, force_stable_argsort=True, backward=True)
        assert msum is None or float(out["mask"].sum()) == msum
        msum = float(out["mask"].sum())
        losses.append(float(out["loss"]))
        frames.append(out["frame_losses"].detach().flatten())
        g = {k: v.double() for k, v in out["grads"].items()}
        if i == 0:
            rec["b1::loss"] = np.float64(losses[0])
            rec["b1::frame_losses"] = frames[0].numpy()
            for k, v in g.items():
                rec["b1::norm::" + k] = np.float64(v.norm())
            for k in GRAD_SLICE_TENSORS:
                rec["b1::slice::" + k] = grad_slice(g[k]).float().numpy()
            acc = g
        else:
            for k in acc:
                acc[k] += g[k]
        print(f"volume {i}: loss {losses[-1]:.6f}", flush=True)
        del out, g
    tag = f"b{batch}::"
    rec[tag + "loss"] = np.float64(sum(losses) / batch)
    rec[tag + "frame_losses"] = torch.stack(frames).numpy()
    for k, v in acc.items():
        v = v / batch
        rec[tag + "norm::" + k] = np.float64(v.norm())
        if k in GRAD_SLICE_TENSORS:
            rec[tag + "slice::" + k] = grad_slice(v).float().numpy()
    rec["mask_sum_per_volume"] = np.float64(msum)
    np.savez_compressed(os.path.join(GOLD, "full_cfg1_grads.npz"), **rec)
    print("full_cfg1_grads.npz:", len(rec), "entries; batch loss", rec[tag + "loss"])


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--full", action="store_true")
    ap.add_argument("--full-grads", action="store_true", help="full_cfg1_grads.npz + full_cfg3_grads.npz (reference fwd+bwd at production size; ~10 min)")
    ap.add_argument("--full-grads-cfg3", action="store_true", help="full_cfg3_grads.npz only (60-frame volume)")
    ap.add_argument("--only-2d", action="store_true", help="regenerate toy2d_step.npz only")
    ap.add_argument("--only-vit", action="store_true", help="regenerate toy_vit_step.npz only")
    a = ap.parse_args()
    if a.full_grads:
        torch.set_num_threads(os.cpu_count())
        gen_full_grads()
        gen_full_grads_cfg3()
        sys.exit(0)
    if a.full_grads_cfg3:
        torch.set_num_threads(os.cpu_count())
        gen_full_grads_cfg3()
        sys.exit(0)
    if a.only_2d:
        gen_toy2d()
        sys.exit(0)
    if a.only_vit:
        gen_toy_vit()
        sys.exit(0)
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    gen_toy(False)
    gen_toy(True)
    gen_masking()
    gen_toy2d()
    gen_toy_vit()
    if a.full:
        gen_full()
