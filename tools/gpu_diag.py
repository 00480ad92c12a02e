"""Verbose per-kernel diagnostics for the B200 box (prints error patterns instead of stopping at the first assert).
Run: python tools/gpu_diag.py [names...]   (output goes to stdout; gpurun copies gpurun_out/)."""
import math
import os
import sys
import time
import traceback

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from octcubem_b200 import _lib, ops  # noqa: E402
from octcubem_b200._lib import (EPI_BIAS, EPI_BIAS_GELU, EPI_DGELU, EPI_NONE, GEMM_NN, GEMM_NT, GEMM_TN, OCT_BF16,  # noqa
                                OCT_F32, OCT_SIMT_BF16)

dev = torch.device("cuda:0")
RESULTS = []


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def report(name, got, want, tol):
    r = rel(got, want)
    finite = bool(torch.isfinite(got.float()).all())
    ok = finite and r <= tol
    RESULTS.append((name, ok, r))
    print(f"[{'OK ' if ok else 'BAD'}] {name}: rel={r:.3e} tol={tol:.1e} finite={finite}", flush=True)
    if not ok and got.dim() == 2:
        d = (got.double().cpu() - want.double().cpu()).abs()
        thr = 1e-2 * float(want.double().abs().max())
        bad = d > thr
        rows = bad.any(1).nonzero().flatten()
        cols = bad.any(0).nonzero().flatten()
        print(f"      bad elems {int(bad.sum())}/{bad.numel()}  bad rows {rows.numel()} (first {rows[:12].tolist()})"
              f"  bad cols {cols.numel()} (first {cols[:12].tolist()})")
        print("      got [0,:8]", got[0, :8].float().cpu().tolist())
        print("      want[0,:8]", want[0, :8].float().cpu().tolist())
    return ok


def t_gemm_tc():
    g = torch.Generator(device="cpu").manual_seed(0)
    shapes = [(128, 256, 64), (128, 128, 64), (256, 512, 256), (384, 768, 512), (3280, 1024, 1024), (200, 136, 72),
              (64, 96, 32), (4097, 512, 512), (1024, 3072, 1024)]
    for layout, lname in ((GEMM_NT, "NT"), (GEMM_NN, "NN"), (GEMM_TN, "TN")):
        for (M, N, K) in shapes:
            a = torch.randn(M, K, generator=g).bfloat16()
            b = torch.randn(N, K, generator=g).bfloat16()
            want = a.float() @ b.float().t()
            A = a.to(dev) if layout != GEMM_TN else a.t().contiguous().to(dev)        # TN: A stored [K, M]
            B = b.to(dev) if layout == GEMM_NT else b.t().contiguous().to(dev)        # NN/TN: B stored [K, N]
            try:
                for od, tol in ((torch.float32, 1e-5), (torch.bfloat16, 4e-3)):
                    out = ops.gemm(layout, A, B, M, N, K, od, compute=OCT_BF16)
                    torch.cuda.synchronize()
                    report(f"gemm_tc {lname} {M}x{N}x{K} -> {str(od)[6:]}", out, want, tol)
            except Exception as e:  # noqa
                print(f"[EXC] gemm_tc {lname} {M}x{N}x{K}: {e}")
                RESULTS.append((f"gemm_tc {lname} {M}x{N}x{K}", False, float('nan')))
                return
    # epilogues
    M, N, K = 384, 512, 256
    a = torch.randn(M, K, generator=g).bfloat16(); b = torch.randn(N, K, generator=g).bfloat16() * 0.1
    bias = torch.randn(N, generator=g)
    pre_ref = (a.float() @ b.float().t() + bias)
    out = ops.gemm(GEMM_NT, a.to(dev), b.to(dev), M, N, K, torch.bfloat16, EPI_BIAS, bias=bias.to(dev), compute=OCT_BF16)
    report("gemm_tc bias", out, pre_ref, 4e-3)
    aux = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
    out = ops.gemm(GEMM_NT, a.to(dev), b.to(dev), M, N, K, torch.bfloat16, EPI_BIAS_GELU, bias=bias.to(dev), aux=aux, compute=OCT_BF16)
    report("gemm_tc bias_gelu aux", aux, pre_ref, 4e-3)
    report("gemm_tc bias_gelu out", out, torch.nn.functional.gelu(pre_ref.bfloat16().float()), 4e-3)
    auxin = torch.randn(M, N, generator=g).bfloat16()
    out = ops.gemm(GEMM_NT, a.to(dev), b.to(dev), M, N, K, torch.bfloat16, EPI_DGELU, aux=auxin.to(dev), compute=OCT_BF16)
    x = auxin.float().requires_grad_(True); torch.nn.functional.gelu(x).sum().backward()
    report("gemm_tc dgelu", out, (a.float() @ b.float().t()) * x.grad, 4e-3)
    acc0 = torch.randn(M, N, generator=g)
    accd = acc0.to(dev).clone()
    ops.gemm(GEMM_NT, a.to(dev), b.to(dev), M, N, K, torch.float32, out=accd, beta=1, compute=OCT_BF16)
    report("gemm_tc beta=1", accd, acc0 + a.float() @ b.float().t(), 1e-5)


def t_gemm_f32():
    g = torch.Generator().manual_seed(1)
    for layout, lname in ((GEMM_NT, "NT"), (GEMM_NN, "NN"), (GEMM_TN, "TN")):
        for (M, N, K) in [(130, 70, 33), (256, 128, 64)]:
            a = torch.randn(M, K, generator=g); b = torch.randn(N, K, generator=g)
            A = a.to(dev) if layout != GEMM_TN else a.t().contiguous().to(dev)
            B = b.to(dev) if layout == GEMM_NT else b.t().contiguous().to(dev)
            out = ops.gemm(layout, A, B, M, N, K, torch.float32, compute=OCT_F32)
            report(f"gemm_f32 {lname} {M}x{N}x{K}", out, a.double() @ b.double().t(), 1e-6)
    M, N, K = 130, 72, 40
    a = torch.randn(M, K, generator=g); b = torch.randn(N, K, generator=g) * 0.2; bias = torch.randn(N, generator=g)
    aux = torch.empty(M, N, device=dev)
    out = ops.gemm(GEMM_NT, a.to(dev), b.to(dev), M, N, K, torch.float32, EPI_BIAS_GELU, bias=bias.to(dev), aux=aux)
    pre = a @ b.t() + bias
    report("gemm_f32 bias_gelu aux", aux, pre, 1e-6)
    report("gemm_f32 bias_gelu out", out, torch.nn.functional.gelu(pre), 1e-6)


def _attn_ref(qkv, H):
    B, S, _ = qkv.shape
    d = qkv.shape[-1] // (3 * H)
    q, k, v = qkv.view(B, S, 3, H, d).unbind(2)
    s = torch.einsum("bthd,bshd->bhts", q, k) / math.sqrt(d)
    p = torch.softmax(s, -1)
    o = torch.einsum("bhts,bshd->bthd", p, v).reshape(B, S, H * d)
    return o, torch.logsumexp(s, -1)


def t_attn(compute, dtype, tol, cases=((2, 77, 2, 32), (1, 300, 2, 64), (2, 130, 1, 16))):
    g = torch.Generator().manual_seed(2)
    for (B, S, H, d) in cases:
        qkv = torch.randn(B, S, 3 * H * d, generator=g).to(dtype)
        dout = torch.randn(B, S, H * d, generator=g).to(dtype)
        x = qkv.double().requires_grad_(True)
        o_ref, lse_ref = _attn_ref(x, H)
        (o_ref * dout.double()).sum().backward()
        try:
            out, lse = ops.attn_fwd(qkv.to(dev), H, d, compute)
            torch.cuda.synchronize()
            report(f"attn_fwd c{compute} B{B} S{S} H{H} d{d} out", out, o_ref.detach(), tol)
            report(f"attn_fwd c{compute} B{B} S{S} H{H} d{d} lse", lse, lse_ref.detach(), 1e-5 if dtype == torch.float32 else 2e-3)
            dq = ops.attn_bwd(qkv.to(dev), out, dout.to(dev), lse, H, d, compute)
            torch.cuda.synchronize()
            gq = x.grad.view(B, S, 3, H, d)
            dq5 = dq.view(B, S, 3, H, d)
            for i, n in enumerate("qkv"):
                report(f"attn_bwd c{compute} B{B} S{S} H{H} d{d} d{n}", dq5[:, :, i].reshape(B * S, -1), gq[:, :, i].reshape(B * S, -1), tol * 2)
        except Exception as e:  # noqa
            print(f"[EXC] attn c{compute} {(B, S, H, d)}: {e}")
            RESULTS.append((f"attn c{compute} {(B, S, H, d)}", False, float('nan')))
            return


def t_patch_embed():
    g = torch.Generator().manual_seed(3)
    for (B, T, HW, E, u) in [(2, 6, 64, 64, 3), (1, 12, 256, 1024, 3), (2, 3, 128, 264, 3), (1, 8, 512, 256, 4)]:
        imgs = torch.rand(B, 1, T, HW, HW, generator=g)
        w = torch.randn(E, 1, u, 16, 16, generator=g) * 0.05
        b = torch.randn(E, generator=g)
        want = torch.nn.functional.conv3d(imgs.double(), w.double(), b.double(), stride=(u, 16, 16)).flatten(2).transpose(1, 2)
        try:
            out = ops.patch_embed_tc(imgs.to(dev), w.view(E, -1).to(dev).contiguous(), b.to(dev), 16, u, torch.float32)
            torch.cuda.synchronize()
            report(f"patch_embed_tc B{B} T{T} {HW}px E{E} u{u}", out.reshape(-1, E), want.reshape(-1, E), 2e-3)
        except Exception as e:  # noqa
            print(f"[EXC] patch_embed {(B, T, HW, E, u)}: {e}")
            RESULTS.append((f"patch_embed {(B, T, HW, E, u)}", False, float('nan')))
            return


def t_pe_probe():
    """One-hot weights: out[tok, e] must equal patch element e of token tok.  Three probes (x, y, frame coordinates)
    decode which input element every output actually reads."""
    B, T, HW, u = 1, 3, 64, 3
    E = 768
    w = torch.eye(768).view(E, 1, 3, 16, 16).contiguous()
    bias = torch.zeros(E)
    f, y, x = torch.meshgrid(torch.arange(T), torch.arange(HW), torch.arange(HW), indexing="ij")
    for name, vol in (("x", x), ("y", y), ("f", f)):
        imgs = vol.float().view(1, 1, T, HW, HW).contiguous()
        out = ops.patch_embed_tc(imgs.to(dev), w.view(E, -1).to(dev).contiguous(), bias.to(dev), 16, u, torch.float32).cpu()
        want = torch.nn.functional.conv3d(imgs, w, bias, stride=(u, 16, 16)).flatten(2).transpose(1, 2)
        print(f"probe {name}: rel {rel(out, want):.3e}")
        for tok in (0, 1, 5):
            print(f"   tok {tok} got  e[0:40]  ", out[0, tok, :40].int().tolist())
            print(f"   tok {tok} want e[0:40]  ", want[0, tok, :40].int().tolist())
            print(f"   tok {tok} got  e[256:272]", out[0, tok, 256:272].int().tolist(), " want", want[0, tok, 256:272].int().tolist())
    # B probe: constant image = 1 -> out[tok, e] = sum_k W[e, k]; use W[e, k] = (k == e % 768) * (e + 1)
    imgs = torch.ones(1, 1, T, HW, HW)
    w2 = torch.zeros(E, 768)
    w2[torch.arange(E), torch.arange(E)] = torch.arange(1, E + 1).float()
    out = ops.patch_embed_tc(imgs.to(dev), w2.to(dev).contiguous(), bias.to(dev), 16, u, torch.float32).cpu()
    print("probe B: got e[0:24]", out[0, 0, :24].int().tolist(), "... e[760:768]", out[0, 0, 760:].int().tolist())


def t_timing():
    """Rough single-kernel timings (L2-warm) for orientation only."""
    def timeit(fn, n=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(n):
            fn()
        e.record(); torch.cuda.synchronize()
        return s.elapsed_time(e) / n
    for (M, N, K) in [(3280, 3072, 1024), (3280, 4096, 1024), (3280, 1024, 4096), (32776, 1536, 512), (32776, 2048, 512),
                      (32776, 512, 2048), (8192, 8192, 8192)]:
        a = torch.randn(M, K, device=dev).bfloat16(); b = torch.randn(N, K, device=dev).bfloat16()
        out = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
        ms = timeit(lambda: ops.gemm(GEMM_NT, a, b, M, N, K, torch.bfloat16, out=out, compute=OCT_BF16))
        ms_t = timeit(lambda: torch.matmul(a, b.t(), out=out))
        print(f"[TIME] gemm NT {M}x{N}x{K}: {ms*1e3:.1f} us = {2*M*N*K/ms/1e9:.0f} TF/s   (torch/cuBLAS {ms_t*1e3:.1f} us = {2*M*N*K/ms_t/1e9:.0f} TF/s)", flush=True)
    for (M, N, K) in [(1024, 4096, 3280), (2048, 512, 32776)]:
        a = torch.randn(K, M, device=dev).bfloat16(); b = torch.randn(K, N, device=dev).bfloat16()
        out = torch.empty(M, N, dtype=torch.float32, device=dev)
        ms = timeit(lambda: ops.gemm(GEMM_TN, a, b, M, N, K, torch.float32, out=out, compute=OCT_BF16))
        print(f"[TIME] gemm TN {M}x{N}x{K}: {ms*1e3:.1f} us = {2*M*N*K/ms/1e9:.0f} TF/s", flush=True)
    imgs = torch.rand(8, 1, 48, 256, 256, device=dev); w = torch.randn(1024, 768, device=dev); bb = torch.randn(1024, device=dev)
    ms = timeit(lambda: ops.patch_embed_tc(imgs, w, bb, 16, 3))
    print(f"[TIME] patch_embed_tc B8 T48: {ms*1e3:.1f} us", flush=True)
    noise = torch.rand(8, 5120, device=dev)
    ms = timeit(lambda: ops.mask_sort(noise, 511))
    print(f"[TIME] mask_sort B8 L5120: {ms*1e3:.1f} us", flush=True)


TESTS = {
    "gemm_f32": t_gemm_f32,
    "gemm_tc": t_gemm_tc,
    "attn_f32": lambda: t_attn(OCT_F32, torch.float32, 2e-5),
    "attn_simt_bf16": lambda: t_attn(OCT_SIMT_BF16, torch.bfloat16, 8e-3),
    "attn_tc": lambda: t_attn(OCT_BF16, torch.bfloat16, 8e-3, cases=((2, 77, 2, 32), (1, 300, 2, 64), (2, 512, 4, 64), (1, 1030, 3, 32))),
    "patch_embed": t_patch_embed,
    "pe_probe": t_pe_probe,
    "timing": t_timing,
}

if __name__ == "__main__":
    names = sys.argv[1:] or list(TESTS)
    print(_lib.version(), torch.cuda.get_device_name(0), flush=True)
    for n in names:
        print(f"=== {n} ===", flush=True)
        t0 = time.time()
        try:
            TESTS[n]()
            torch.cuda.synchronize()
        except Exception:
            traceback.print_exc()
            RESULTS.append((n, False, float("nan")))
        print(f"--- {n}: {time.time() - t0:.1f}s", flush=True)
    bad = [r for r in RESULTS if not r[1]]
    print(f"SUMMARY: {len(RESULTS) - len(bad)} ok, {len(bad)} bad")
    for r in bad:
        print("  BAD:", r[0], r[2])
