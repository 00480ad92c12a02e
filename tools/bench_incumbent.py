"""Kernel by kernel against the stock stack the reference runs on a GPU (SURVEY §2.2 'bar:' lines), same box, same shapes,
L2 flushed between timed launches, median of 10:
    every GEMM of the step        ours (tcgen05)            vs  F.linear / torch.matmul (cuBLASLt), + F.gelu where ours fuses it
    attention fwd / bwd           ours (tcgen05 flash)      vs  flash_attn_qkvpacked_func (FA2 2.8.3)  [and torch SDPA]
    patch embedding               ours (TMA im2col-free)    vs  F.conv3d + permute (cuDNN)
    add + LayerNorm               ours                      vs  add + F.layer_norm (ATen)
    step                          bench.py's own numbers    vs  oracle/gpu_incumbent.py (eager and CUDA graph)
Writes a markdown table (default profiles/r2_incumbent.md) and a JSON next to it.  Measurement tool, not product code.

    python tools/bench_incumbent.py [--out profiles/r2_incumbent.md] [--skip-step]
"""
import argparse
import json
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from octcubem_b200 import ops  # noqa: E402
from octcubem_b200._lib import EPI_BIAS, EPI_BIAS_GELU, EPI_DGELU, EPI_NONE, GEMM_NN, GEMM_NT, OCT_BF16  # noqa: E402

dev = torch.device("cuda:0")
flush = None


def med_us(fn, n=10, warm=3):
    global flush
    if flush is None:
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def gemm_rows(frames):
    Te = 8 * (int((frames // 3) * 256 * (1 - 0.9)) + 1)
    Td = 8 * ((frames // 3) * 256 + 1)
    rows = []
    for tag, T, C, reps in (("enc", Te, 1024, 24), ("dec", Td, 512, 8)):
        for name, kind, M, N, K, epi in ((f"{tag} Wqkv fwd", "NT", T, 3 * C, C, EPI_BIAS), (f"{tag} out_proj fwd", "NT", T, C, C, EPI_BIAS),
                                         (f"{tag} fc1 fwd + GELU", "NT", T, 4 * C, C, EPI_BIAS_GELU), (f"{tag} fc2 fwd", "NT", T, C, 4 * C, EPI_BIAS),
                                         (f"{tag} Wqkv dgrad", "NN", T, C, 3 * C, EPI_NONE), (f"{tag} out_proj dgrad", "NN", T, C, C, EPI_NONE),
                                         (f"{tag} fc1 dgrad", "NN", T, C, 4 * C, EPI_NONE), (f"{tag} fc2 dgrad + dGELU", "NN", T, 4 * C, C, EPI_DGELU),
                                         (f"{tag} Wqkv wgrad + db", "WG", 3 * C, C, T, 0), (f"{tag} fc1 wgrad + db", "WG", 4 * C, C, T, 0),
                                         (f"{tag} fc2 wgrad + db", "WG", C, 4 * C, T, 0), (f"{tag} out_proj wgrad + db", "WG", C, C, T, 0)):
            if kind == "WG":
                dy = torch.randn(K, M, device=dev).bfloat16(); x = torch.randn(K, N, device=dev).bfloat16()
                ours = med_us(lambda: ops.wgrad_bias(dy, x))
                lib = med_us(lambda: (torch.matmul(dy.t(), x), dy.sum(0)))           # autograd's Linear backward: mm + sum
                note = "cuBLASLt mm (bf16 out) + reduce"
            else:
                a = torch.randn(M, K, device=dev).bfloat16()
                w = torch.randn((N, K) if kind == "NT" else (K, N), device=dev).bfloat16()
                bias = torch.randn(N, device=dev)
                bias_lp = bias.bfloat16()
                out = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
                aux = torch.randn(M, N, device=dev).bfloat16() if epi in (EPI_BIAS_GELU, EPI_DGELU) else None
                lay = GEMM_NT if kind == "NT" else GEMM_NN
                ours = med_us(lambda: ops.gemm(lay, a, w, M, N, K, torch.bfloat16, epi, bias=bias if epi in (EPI_BIAS, EPI_BIAS_GELU) else None,
                                               aux=aux, out=out, compute=OCT_BF16))
                if epi == EPI_BIAS:
                    lib, note = med_us(lambda: F.linear(a, w, bias_lp)), "F.linear"
                elif epi == EPI_BIAS_GELU:
                    lib, note = med_us(lambda: F.gelu(F.linear(a, w, bias_lp))), "F.linear + F.gelu"
                elif epi == EPI_DGELU:
                    pre = aux.clone().requires_grad_(True)
                    lib = med_us(lambda: torch.ops.aten.gelu_backward(torch.matmul(a, w), pre))
                    note = "matmul + gelu_backward"
                else:
                    lib, note = med_us(lambda: torch.matmul(a, w)), "torch.matmul"
            fl = 2.0 * M * N * K
            rows.append({"kernel": name, "shape": [M, N, K], "launches_per_step": reps, "ours_us": ours, "lib_us": lib, "lib": note,
                         "ours_tflops": fl / ours / 1e6, "lib_tflops": fl / lib / 1e6})
            print(rows[-1], flush=True)
    return rows


def attention_rows(frames):
    rows = []
    try:
        from flash_attn import flash_attn_qkvpacked_func
    except Exception as e:  # noqa: BLE001
        flash_attn_qkvpacked_func = None
        print("flash_attn import failed:", e)
    Sd = (frames // 3) * 256 + 1
    Se = int((Sd - 1) * (1 - 0.9)) + 1
    for tag, B, S, H, d, reps in (("dec", 8, Sd, 16, 32, 8), ("enc", 8, Se, 16, 64, 24)):
        qkv = (torch.randn(B, S, 3 * H * d, device=dev) * 0.5).bfloat16()
        dout = torch.randn(B, S, H * d, device=dev).bfloat16()
        out, lse = ops.attn_fwd(qkv, H, d, OCT_BF16)
        of = med_us(lambda: ops.attn_fwd(qkv, H, d, OCT_BF16))
        ob = med_us(lambda: ops.attn_bwd(qkv, out, dout, lse, H, d, OCT_BF16))
        rec = {"kernel": f"{tag} attention", "shape": [B, S, H, d], "launches_per_step": reps, "ours_fwd_us": of, "ours_bwd_us": ob}
        q5 = qkv.view(B, S, 3, H, d).clone().requires_grad_(True)
        do4 = dout.view(B, S, H, d)
        if flash_attn_qkvpacked_func is not None:
            try:
                ff = med_us(lambda: flash_attn_qkvpacked_func(q5, 0.0, softmax_scale=d ** -0.5, causal=False))

                def fb():
                    q5.grad = None
                    flash_attn_qkvpacked_func(q5, 0.0, softmax_scale=d ** -0.5, causal=False).backward(do4)
                rec["fa2_fwd_us"], rec["fa2_bwd_us"] = ff, med_us(fb) - ff
            except Exception as e:  # noqa: BLE001
                rec["fa2"] = f"{type(e).__name__}: {e}"[:200]
        try:
            qh, kh, vh = [t.transpose(1, 2) for t in q5.unbind(2)]
            sf = med_us(lambda: F.scaled_dot_product_attention(qh, kh, vh))

            def sb():
                q5.grad = None
                F.scaled_dot_product_attention(qh, kh, vh).backward(do4.transpose(1, 2))
            rec["sdpa_fwd_us"], rec["sdpa_bwd_us"] = sf, med_us(sb) - sf
        except Exception as e:  # noqa: BLE001
            rec["sdpa"] = f"{type(e).__name__}: {e}"[:200]
        rows.append(rec)
        print(rec, flush=True)
    return rows


def misc_rows(frames):
    rows = []
    B, E = 8, 1024
    vol = torch.rand(B, 1, frames, 256, 256, device=dev)
    w = torch.randn(E, 1, 3, 16, 16, device=dev) * 0.02
    b = torch.randn(E, device=dev)
    w2 = w.view(E, -1).contiguous()
    ours = med_us(lambda: ops.patch_embed_tc(vol, w2, b, 16, 3, torch.bfloat16))

    def conv():
        with torch.autocast("cuda", dtype=torch.bfloat16):
            x = F.conv3d(vol, w, b, stride=(3, 16, 16)).flatten(3)
            return torch.einsum("ncts->ntsc", x).contiguous()
    rows.append({"kernel": "patch embedding (dense)", "shape": list(vol.shape), "launches_per_step": 1, "ours_us": ours, "lib_us": med_us(conv),
                 "lib": "autocast F.conv3d (cuDNN) + permute copy"})
    print(rows[-1], flush=True)
    for tag, M, C, reps in (("enc", 8 * (int((frames // 3) * 256 * 0.1) + 1), 1024, 48), ("dec", 8 * ((frames // 3) * 256 + 1), 512, 16)):
        h = torch.randn(M, C, device=dev).bfloat16()
        res = torch.randn(M, C, device=dev)
        g, bb = torch.ones(C, device=dev), torch.zeros(C, device=dev)
        ours = med_us(lambda: ops.add_ln_fwd(h, res, g, bb, 1e-6, torch.bfloat16, True))

        def ln():
            r = h + res                                                          # flash_attn Block: dropped + residual (fp32)
            return F.layer_norm(r, (C,), g, bb, 1e-6).bfloat16(), r
        rows.append({"kernel": f"{tag} add + LayerNorm fwd", "shape": [M, C], "launches_per_step": reps, "ours_us": ours, "lib_us": med_us(ln),
                     "lib": "ATen add + layer_norm + cast"})
        print(rows[-1], flush=True)
    return rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r2_incumbent.md"))
    ap.add_argument("--frames", type=int, default=48)
    ap.add_argument("--skip-step", action="store_true")
    a = ap.parse_args()
    rec = {"frames": a.frames, "gpu": torch.cuda.get_device_name(0), "torch": torch.__version__}
    try:
        import flash_attn
        rec["flash_attn"] = flash_attn.__version__
    except Exception as e:  # noqa: BLE001
        rec["flash_attn"] = f"unavailable: {e}"
    rec["gemm"] = gemm_rows(a.frames)
    rec["attention"] = attention_rows(a.frames)
    rec["misc"] = misc_rows(a.frames)
    if not a.skip_step:
        from oracle import gpu_incumbent as G

        def timed(fn, n):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); s.record()
            for _ in range(n):
                fn()
            e.record(); torch.cuda.synchronize()
            return s.elapsed_time(e)
        rec["step"] = G.bench_block(a.frames, 256, 8, 0.9, dev, timed, 10)
        print(rec["step"], flush=True)
    json.dump(rec, open(os.path.splitext(a.out)[0] + ".json", "w"), indent=1)
    L = [f"# Same-box incumbents, kernel by kernel (round 2) — {rec['gpu']}, torch {rec['torch']}, flash_attn {rec['flash_attn']}", "",
         f"`python tools/bench_incumbent.py --frames {a.frames}`: B = 8 volumes of {a.frames}x256x256 (BASELINE configs[{1 if a.frames == 48 else 2}]), bf16, L2 flushed between "
         "timed launches, median of 10 (CUDA events).  'lib' = the stock kernel the reference's GPU path runs (SURVEY §2.2).", "",
         "| GEMM | M x N x K | per step | ours us | ours TF/s | lib us | lib TF/s | lib / ours | lib kernel(s) |", "|---|---|---:|---:|---:|---:|---:|---:|---|"]
    to = tl = 0.0
    for r in rec["gemm"]:
        to += r["ours_us"] * r["launches_per_step"]; tl += r["lib_us"] * r["launches_per_step"]
        L.append(f"| {r['kernel']} | {'x'.join(map(str, r['shape']))} | {r['launches_per_step']} | {r['ours_us']:.1f} | {r['ours_tflops']:.0f} | "
                 f"{r['lib_us']:.1f} | {r['lib_tflops']:.0f} | {r['lib_us'] / r['ours_us']:.2f} | {r['lib']} |")
    L += [f"| **all GEMMs of one step** | | | **{to / 1e3:.2f} ms** | | **{tl / 1e3:.2f} ms** | | **{tl / to:.2f}** | |", "",
          "| attention | B,S,H,d | per step | ours fwd us | ours bwd us | FA2 fwd us | FA2 bwd us | SDPA fwd us | SDPA bwd us | FA2/ours fwd | FA2/ours bwd |",
          "|---|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|"]
    for r in rec["attention"]:
        f = lambda k: f"{r[k]:.1f}" if k in r else "n/a"  # noqa: E731
        rf = f"{r['fa2_fwd_us'] / r['ours_fwd_us']:.2f}" if "fa2_fwd_us" in r else "n/a"
        rb = f"{r['fa2_bwd_us'] / r['ours_bwd_us']:.2f}" if "fa2_bwd_us" in r else "n/a"
        L.append(f"| {r['kernel']} | {r['shape']} | {r['launches_per_step']} | {f('ours_fwd_us')} | {f('ours_bwd_us')} | {f('fa2_fwd_us')} | {f('fa2_bwd_us')} | "
                 f"{f('sdpa_fwd_us')} | {f('sdpa_bwd_us')} | {rf} | {rb} |")
        for k in ("fa2", "sdpa"):
            if k in r:
                L.append(f"| ({k}: {r[k]}) | | | | | | | | | | |")
    L += ["", "| other | shape | per step | ours us | lib us | lib / ours | lib kernel(s) |", "|---|---|---:|---:|---:|---:|---|"]
    for r in rec["misc"]:
        L.append(f"| {r['kernel']} | {r['shape']} | {r['launches_per_step']} | {r['ours_us']:.1f} | {r['lib_us']:.1f} | {r['lib_us'] / r['ours_us']:.2f} | {r['lib']} |")
    if "step" in rec:
        s = rec["step"]
        L += ["", "## Whole step (forward + backward, batch 8)", "", "```", json.dumps(s, indent=1), "```"]
    open(a.out, "w").write("\n".join(L) + "\n")
    print("wrote", a.out)


if __name__ == "__main__":
    main()
