"""Import harness for the UNMODIFIED reference 3D MAE (test infrastructure, not product code).

Only usable where /root/reference exists (the build container).  Nothing under tests -m gpu,
smoke() or bench.py imports this file.  It is used to (a) pin oracle/mae3d_oracle.py against the
real reference and (b) generate the committed fixtures under tests/golden/ (oracle/gen_golden.py).

Recipe follows SURVEY.md §8(c) / Appendix A: stub the absent third-party imports (timm, iopath,
simplejson), import Pre-training/models_mae_joint_res_flash_attn.py as-is, build the flash variant
and swap `mixer.inner_attn` for flash_attn.modules.mha.SelfAttention (pure torch) so it runs in
fp32 on CPU with the very same parameters / Block control flow (incl. quirk Q1).
"""
import builtins
import contextlib
import json
import os
import sys
import types

import torch
import torch.nn as nn

REF_ROOT = os.environ.get("OCT_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "Pre-training", "models_mae_joint_res_flash_attn.py"))


def _install_stubs():
    def mod(name, **attrs):
        m = sys.modules.get(name)
        if m is None:
            m = types.ModuleType(name)
            m.__path__ = []  # behave as a package
            sys.modules[name] = m
        for k, v in attrs.items():
            setattr(m, k, v)
        return m

    def to_2tuple(x):
        return tuple(x) if isinstance(x, (tuple, list)) else (x, x)

    class DropPath(nn.Module):
        def __init__(self, p=0.0):
            super().__init__()
            self.p = p

        def forward(self, x):
            return x

    class Mlp(nn.Module):
        def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.0):
            super().__init__()
            out_features = out_features or in_features
            hidden_features = hidden_features or in_features
            self.fc1 = nn.Linear(in_features, hidden_features)
            self.act = act_layer()
            self.fc2 = nn.Linear(hidden_features, out_features)

        def forward(self, x):
            return self.fc2(self.act(self.fc1(x)))

    if getattr(sys.modules.get("timm"), "__file__", None) is None:  # absent, or a stub (ours / oracle.gpu_incumbent's): complete it
        mod("timm")
        mod("timm.models")
        mod("timm.layers", to_2tuple=to_2tuple)
        mod("timm.models.layers", to_2tuple=to_2tuple)
        mod("timm.models.vision_transformer", DropPath=DropPath, Mlp=Mlp)
        mod("timm.models.helpers", named_apply=lambda *a, **k: None)
    if "iopath" not in sys.modules:
        mod("iopath")
        mod("iopath.common")
        mod("iopath.common.file_io", g_pathmgr=types.SimpleNamespace(open=builtins.open))
    if "simplejson" not in sys.modules:
        sys.modules["simplejson"] = json
    # the 2D twin (OCTCube/models_mae_flash_attn.py) imports timm's Block by name and util/misc.py, which pulls plotting
    # packages in at import time; none of it is on the step path
    vt = sys.modules["timm.models.vision_transformer"]
    if not hasattr(vt, "Block"):
        vt.Block = type("Block", (nn.Module,), {})
    for name in ("matplotlib", "matplotlib.pyplot", "PIL", "PIL.Image", "torchvision", "torchvision.transforms"):
        try:
            __import__(name)
        except Exception:
            parent, _, leaf = name.rpartition(".")
            m = mod(name)
            if parent:
                setattr(sys.modules[parent], leaf, m)
    if not torch.cuda.is_available():
        # flash_attn/ops/triton/layer_norm.py touches torch.cuda at import time
        sys.modules.setdefault("flash_attn.ops.triton.layer_norm", None)


_REF_MODULE = None


def import_reference():
    """Returns the reference module models_mae_joint_res_flash_attn (unmodified source)."""
    global _REF_MODULE
    if _REF_MODULE is None:
        if not reference_available():
            raise RuntimeError(f"reference tree not found under {REF_ROOT}")
        _install_stubs()
        p = os.path.join(REF_ROOT, "Pre-training")
        if p not in sys.path:
            sys.path.insert(0, p)
        import models_mae_joint_res_flash_attn as M  # noqa

        _REF_MODULE = M
    return _REF_MODULE


def build_reference(flash_semantics=True, seed=0, **kw):
    """Builds the reference MaskedAutoencoderViT on CPU fp32.

    flash_semantics=True : create_block blocks with inner_attn -> SelfAttention (the parity oracle object)
    flash_semantics=False: use_flash_attn=False (video_vit.Block; the literal CPU baseline of BASELINE.md §4)
    """
    M = import_reference()
    from functools import partial

    torch.manual_seed(seed)
    with contextlib.redirect_stdout(open(os.devnull, "w")):
        m = M.MaskedAutoencoderViT(
            norm_layer=partial(nn.LayerNorm, eps=1e-6), use_flash_attn=flash_semantics, **kw
        )
    if flash_semantics:
        from flash_attn.modules.mha import SelfAttention

        for blk in list(m.blocks) + list(m.decoder_blocks):
            blk.mixer.inner_attn = SelfAttention()
            blk.mixer.use_flash_attn = False
    return m


_REF_MODULE_2D = None


def import_reference_2d():
    """Returns the reference module OCTCube/models_mae_flash_attn (unmodified source; the 2D twin of SURVEY §8a)."""
    global _REF_MODULE_2D
    if _REF_MODULE_2D is None:
        path = os.path.join(REF_ROOT, "OCTCube", "models_mae_flash_attn.py")
        if not os.path.isfile(path):
            raise RuntimeError(f"reference tree not found under {REF_ROOT}")
        _install_stubs()
        import importlib.util

        p = os.path.join(REF_ROOT, "OCTCube")
        if p not in sys.path:
            sys.path.append(p)  # its `from util.… import …` fallbacks resolve against OCTCube/util
        spec = importlib.util.spec_from_file_location("octcube_ref_models_mae_flash_attn_2d", path)
        M = importlib.util.module_from_spec(spec)
        with contextlib.redirect_stdout(open(os.devnull, "w")):
            spec.loader.exec_module(M)
        _REF_MODULE_2D = M
    return _REF_MODULE_2D


def build_reference_2d(seed=0, **kw):
    """The reference's 2D MaskedAutoencoderViT (flash blocks, SelfAttention swapped in) on CPU fp32."""
    M = import_reference_2d()
    from functools import partial

    from flash_attn.modules.mha import SelfAttention

    torch.manual_seed(seed)
    with contextlib.redirect_stdout(open(os.devnull, "w")):
        m = M.MaskedAutoencoderViT(norm_layer=partial(nn.LayerNorm, eps=1e-6), use_flash_attn=True, **kw)
    for blk in list(m.blocks) + list(m.decoder_blocks):
        blk.mixer.inner_attn = SelfAttention()
        blk.mixer.use_flash_attn = False
    return m


def run_reference_2d(m, imgs, noise, mask_ratio, force_stable_argsort=False, backward=False):
    with contextlib.redirect_stdout(open(os.devnull, "w")), inject_noise(noise, force_stable_argsort):
        loss, pred, mask, frame_loss = m(imgs, mask_ratio=mask_ratio, return_frame_loss=True)
    out = {"loss": loss, "pred": pred, "mask": mask, "frame_loss": frame_loss}
    if backward:
        m.zero_grad(set_to_none=True)
        loss.backward()
        out["grads"] = {k: p.grad.detach().clone() for k, p in m.named_parameters() if p.grad is not None}
    return out


_REF_MODULE_VIT = None


def import_reference_vit():
    """Returns the reference module OCTCube/models_vit_st_flash_attn (unmodified source)."""
    global _REF_MODULE_VIT
    if _REF_MODULE_VIT is None:
        path = os.path.join(REF_ROOT, "OCTCube", "models_vit_st_flash_attn.py")
        if not os.path.isfile(path):
            raise RuntimeError(f"reference tree not found under {REF_ROOT}")
        _install_stubs()
        import importlib.util

        p = os.path.join(REF_ROOT, "OCTCube")
        if p not in sys.path:
            sys.path.append(p)
        spec = importlib.util.spec_from_file_location("octcube_ref_models_vit_st_flash_attn", path)
        M = importlib.util.module_from_spec(spec)
        with contextlib.redirect_stdout(open(os.devnull, "w")):
            spec.loader.exec_module(M)
        _REF_MODULE_VIT = M
    return _REF_MODULE_VIT


def build_reference_vit(seed=0, **kw):
    """The reference's encoder-only VisionTransformer (flash blocks, SelfAttention swapped in) on CPU fp32, eval mode."""
    M = import_reference_vit()
    from functools import partial

    from flash_attn.modules.mha import SelfAttention

    torch.manual_seed(seed)
    with contextlib.redirect_stdout(open(os.devnull, "w")):
        m = M.VisionTransformer(norm_layer=partial(nn.LayerNorm, eps=1e-6), use_flash_attn=True, **kw)
    for blk in m.blocks:
        blk.mixer.inner_attn = SelfAttention()
        blk.mixer.use_flash_attn = False
    return m.eval()


@contextlib.contextmanager
def inject_noise(noise, force_stable_argsort=False):
    """Replaces torch.rand(N, L, device=...) (the only call shape in random_masking,
    models_mae_joint_res_flash_attn.py:350) by the given tensor.  Optionally forces stable argsort
    to mimic the CUDA radix sort the reference really runs on."""
    real_rand, real_argsort = torch.rand, torch.argsort

    def fake_rand(*size, **kw):
        if len(size) == 2 and tuple(size) == tuple(noise.shape):
            return noise.clone()
        return real_rand(*size, **kw)

    def stable_argsort(x, dim=-1, descending=False, stable=False):
        return real_argsort(x, dim=dim, descending=descending, stable=True)

    torch.rand = fake_rand
    if force_stable_argsort:
        torch.argsort = stable_argsort
    try:
        yield
    finally:
        torch.rand, torch.argsort = real_rand, real_argsort


def run_reference(m, imgs, noise, mask_ratio, frame_loss=False, force_stable_argsort=False, backward=False):
    with contextlib.redirect_stdout(open(os.devnull, "w")), inject_noise(noise, force_stable_argsort):
        loss, pred, mask = m(imgs, mask_ratio=mask_ratio, frame_loss=frame_loss)
    out = {"pred": pred, "mask": mask}
    if frame_loss:
        out["loss"], out["frame_losses"] = loss
    else:
        out["loss"] = loss
    if backward:
        m.zero_grad(set_to_none=True)
        out["loss"].backward()
        out["grads"] = {k: p.grad.detach().clone() for k, p in m.named_parameters() if p.grad is not None}
    return out
