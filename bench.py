#!/usr/bin/env python
"""bench.py — 3D-MAE pre-training step throughput (volumes/s) on N B200s of one node.

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
           bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference algorithm's CPU path on this box's host cores

Workload (BASELINE.json configs[1]): MaskedAutoencoderViT ViT-L encoder / 512x8x16 decoder, patch 16, t_patch 3,
48x256x256 single-channel volumes, mask 0.9, bf16 forward+backward, batch 8 per GPU, synthetic volumes, random-init
weights.  A "step" = forward + backward (+ gradient all-reduce over NCCL when N > 1); the optimizer step is NOT part of
the hot path and is reported separately (`optimizer_ms`, torch fused AdamW, informational).
Prints ONE JSON line (see the README / DESIGN.md §Measurement for the fields).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FRAMES, IMG, BATCH, MASK = 48, 256, 8, 0.9
# algorithmic FLOPs per volume, fwd+bwd: 2MNK per GEMM, 4 S^2 dim per attention layer, bwd = 2x fwd
GF_PER_VOLUME = {48: 2260.0, 60: 3097.0}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        d = json.load(open(path))
        return {"bf16_burst": d["bf16_tflops"], "bf16_sustained": d["bf16_tflops_sustained"], "hbm": d["hbm_gbs"],
                "src": "measured (MEASURED_PEAKS.json)"}
    return {"bf16_burst": 1590.0, "bf16_sustained": 1400.0, "hbm": 6650.0, "src": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region (recipe of B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference algorithm's CPU path (models_mae.py-style non-flash blocks, fp32)
# ----------------------------------------------------------------------------------------------------------------
def cpu_reference_step_fn():
    """Returns (fn, description): fn() runs forward+backward of ONE 48x256x256 volume through oracle.cpu_baseline_forward
    (the restatement of the reference class with use_flash_attn=False, pinned against the reference in tests/)."""
    from oracle import mae3d_oracle as O
    cfg = O.MAEConfig(num_frames=FRAMES, pred_t_dim=FRAMES)
    torch.manual_seed(0)
    sd = {k: v.requires_grad_(True) for k, v in O.init_state_dict(cfg, seed=0).items()}
    vol = O.synthetic_volume(1, FRAMES, IMG, IMG, seed=0)
    noise = O.synthetic_noise(1, cfg.t_grid * cfg.grid ** 2, seed=1)

    def fn():
        for v in sd.values():
            v.grad = None
        loss, _, _ = O.cpu_baseline_forward(cfg, sd, vol, MASK, noise)
        loss.backward()
        return float(loss.detach())

    return fn, "1 volume (of the batch-8 workload) forward+backward, fp32, torch CPU kernels"


def run_cpu(max_steps, warmup, budget_s):
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    fn, sample = cpu_reference_step_fn()
    t0 = time.time()
    fn()                                   # first call doubles as warm-up (allocator, thread pool)
    first = time.time() - t0
    times = []
    if first * 2 < budget_s:               # room for at least one more: discard the warm-up call
        for _ in range(max(0, min(warmup, 1) - 1)):
            fn()
        while len(times) < max_steps and (time.time() - t0) + (times[-1] if times else first) < budget_s:
            t1 = time.time(); fn(); times.append(time.time() - t1)
    if not times:
        times = [first]
    times.sort()
    med = times[len(times) // 2]
    return {"value": 1.0 / med, "unit": "volumes/s", "cores": cores, "kind": "port",
            "sample": f"{sample}; {len(times)} timed run(s), median {med:.2f} s/volume", "steps_run": len(times),
            "ms_per_step": med * 1e3}


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = run_cpu(args.steps, args.warmup, budget_s=150.0)
    line = {"metric": "3D-MAE pretrain volumes/sec", "value": r["value"], "unit": "volumes/s", "n_gpus": args.gpus,
            "steps": r["steps_run"], "warmup": min(args.warmup, 1), "ms_per_step": r["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "impl": "reference",
            "config": {"workload": workload_name(), "sample": r["sample"]},
            "cpu_baseline": {"value": r["value"], "unit": "volumes/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]},
            "e2e": {"value": r["value"], "unit": "volumes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_name(frames=None):
    frames = FRAMES if frames is None else frames
    return (f"3D MAE ViT-L/16 t_patch 3 + dec 512x8x16, {frames}x{IMG}x{IMG} volumes, mask {MASK}, bf16 fwd+bwd, "
            f"batch {BATCH}/GPU (BASELINE.json configs[{1 if frames == 48 else 2}])")


# ----------------------------------------------------------------------------------------------------------------
# B200 arm
# ----------------------------------------------------------------------------------------------------------------
def _watchdog(seconds):
    time.sleep(seconds)
    print(f"[bench] watchdog: still running after {seconds} s, aborting", file=sys.stderr, flush=True)
    os._exit(3)


class StepBench:
    """One configuration (frame count) of the step on this rank: model, optional reducer, eager warm-up, CUDA graph."""

    def __init__(self, frames, args, dev, world, rank, lib, use_reducer=True):
        import torch.distributed as dist
        from octcubem_b200 import models_mae
        from octcubem_b200.dp import GradReducer
        self.frames, self.dev, self.world = frames, dev, world
        torch.manual_seed(1234)
        model = models_mae.flash_attn_mae_vit_large_patch16(
            input_size=IMG, in_chans=1, num_frames=frames, t_patch_size=3, pred_t_dim=frames, sep_pos_embed=True,
            cls_embed=True, high_res_input_size=512, decoder_embed_dim=512, decoder_depth=8, decoder_num_heads=16,
            precision="bf16").to(dev)
        if world > 1:  # same weights on every rank (DDP broadcasts rank 0's at construction)
            for p in model.parameters():
                dist.broadcast(p.data, 0)
        bucket_mb = float(os.environ.get("OCT_BUCKET_MB", "32"))  # experiment knob; 32 MB is the measured default
        self.model = model
        self.reducer = GradReducer(model, bucket_mb=bucket_mb) if (world > 1 and use_reducer) else None
        g = torch.Generator().manual_seed(100 + rank)
        host_vol = torch.rand(BATCH, 1, frames, IMG, IMG, generator=g)
        host_vol[:, :, :3] = 0; host_vol[:, :, -3:] = 0       # centre-padding of PatientDataset_inhouse.py:439-444
        self.host_vol = host_vol.pin_memory()
        self.vol = self.host_vol.to(dev)
        self.loss_out = torch.zeros((), device=dev)
        # eager warm-up (also: TMA maps / func attributes / reducer bucket discovery), launch count of one step
        for _ in range(2):
            self.step()
        torch.cuda.synchronize()
        c0 = lib.oct_launch_count()
        self.step()
        torch.cuda.synchronize()
        self.launches_per_step = int(lib.oct_launch_count() - c0)
        self.graph = None
        if not args.no_graph:
            try:
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    self.step()
                torch.cuda.current_stream().wait_stream(side)
                torch.cuda.synchronize()
                self.graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.graph):
                    self.step()
                self.graph.replay()
                torch.cuda.synchronize()
            except Exception as e:  # noqa
                if rank == 0:
                    print(f"[bench] CUDA graph capture failed ({type(e).__name__}: {e}); running eagerly", file=sys.stderr)
                self.graph = None
                torch.cuda.synchronize()

    def step(self):
        model, reducer = self.model, self.reducer
        if reducer is not None:
            reducer.zero_grad()
        else:
            model.zero_grad(set_to_none=True)
        loss, _, _ = model(self.vol, mask_ratio=MASK)           # noise drawn on device, like models...:350
        if reducer is not None:
            reducer.backward(loss)                              # seeded with 1/world: the all-reduce yields DDP's mean
            reducer.finish()
        else:
            loss.backward()
        self.loss_out.copy_(loss.detach())

    def run(self):
        if self.graph is not None:
            self.graph.replay()
        else:
            self.step()

    def close(self):
        """Drop the captured graph (it holds NCCL kernels of the process group) and the reducer's hooks / buckets."""
        self.graph = None
        if self.reducer is not None:
            self.reducer.remove()
            self.reducer = None
        self.model = None
        self.vol = self.host_vol = None
        import gc
        gc.collect()
        torch.cuda.synchronize()
        torch.cuda.empty_cache()


def main_b200(args):
    import torch.distributed as dist
    from octcubem_b200 import _lib, ops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
        os.environ.pop("NCCL_DEBUG")  # NCCL prints its version banner on stdout at these levels; keep stdout = one JSON line
    threading.Thread(target=_watchdog, args=(1500,), daemon=True).start()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    pk = peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(n):
            fn()
        e.record()
        torch.cuda.synchronize()
        ms = torch.tensor([s.elapsed_time(e)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        barrier()
        return float(ms)

    sb = StepBench(FRAMES, args, dev, world, rank, lib)
    model, vol, host_vol, loss_out = sb.model, sb.vol, sb.host_vol, sb.loss_out
    launches_per_step, run = sb.launches_per_step, sb.run
    graph_used = sb.graph is not None
    for _ in range(max(args.warmup, 3)):
        run()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    ms_total = timed(run, args.steps)
    clocks = sampler.stop() if sampler else None
    ms_step = ms_total / args.steps
    value = world * BATCH / (ms_step / 1e3)

    # end to end through the public API: pinned host volume -> device, step, loss -> host, every step.  Like any training
    # input pipeline (pin_memory + non_blocking prefetch) the upload of step i+1 runs on a copy stream while step i
    # computes: H2D lands in a staging buffer, a device-side copy moves it into the graph's input at the top of the step.
    copy_stream = torch.cuda.Stream()
    stage = torch.empty_like(vol)
    h2d_done, stage_free = torch.cuda.Event(), torch.cuda.Event()

    def upload():
        copy_stream.wait_event(stage_free)
        with torch.cuda.stream(copy_stream):
            stage.copy_(host_vol, non_blocking=True)
            h2d_done.record(copy_stream)

    stage_free.record()
    upload()

    def e2e_step():
        cur = torch.cuda.current_stream()
        cur.wait_event(h2d_done)
        vol.copy_(stage, non_blocking=True)
        stage_free.record(cur)
        upload()                      # next step's volumes: overlaps this step's compute
        run()
        return loss_out.item()        # device -> host read of the loss: the host observes every step's result

    for _ in range(3):
        e2e_step()
    e2e_ms = timed(e2e_step, args.steps) / args.steps
    e2e_value = world * BATCH / (e2e_ms / 1e3)
    last_loss = float(loss_out)
    torch.cuda.synchronize()
    del stage

    # N > 1: the same step WITHOUT the gradient exchange (local gradients only) on the same build, same box: the difference
    # to the step above is the communication that backward does not hide (plus NCCL's SM footprint while it overlaps)
    comm = None
    if world > 1 and not args.no_extras:
        try:
            used_graph = sb.graph is not None
            sb.graph = None
            sb.reducer.remove()
            sb.reducer = None
            torch.cuda.synchronize()
            for _ in range(2):
                sb.step()
            torch.cuda.synchronize()
            g2 = None
            if used_graph:
                g2 = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g2):
                    sb.step()
            local_run = g2.replay if g2 is not None else sb.step
            for _ in range(3):
                local_run()
            ms_local = timed(local_run, args.steps) / args.steps
            comm = {"step_ms_without_exchange": ms_local, "exposed_comm_ms": ms_step - ms_local,
                    "note": "same ranks, same build; the reducer removed, gradients stay local"}
            del g2
        except Exception as e:  # noqa
            if rank == 0:
                print(f"[bench] no-exchange timing skipped ({type(e).__name__}: {e})", file=sys.stderr)

    # roofline of the dominant kernel, timed alone at its in-step shape (CUDA events on the launching stream)
    roof = dominant_kernel_roofline(ops, dev, pk)

    # optimizer step, informational (not on the hot path): our fused multi-tensor AdamW (one launch per
    # parameter group, also emits the bf16 weight shadows) next to torch's fused AdamW on the same parameters
    opt_ms = opt_torch_ms = None
    if not args.no_extras:
        try:
            from octcubem_b200 import optim
            if sb.reducer is None and world > 1:
                sb.step()                                           # plain gradients again after the reducer was removed
            opt = optim.FusedAdamW(optim.add_weight_decay(model, 0.05), lr=1e-6, betas=(0.9, 0.95), shadows=model.shadow_of)
            for _ in range(2):
                opt.step()
            opt_ms = timed(opt.step, 5) / 5
            model.shadows_current()
            topt = torch.optim.AdamW(optim.add_weight_decay(model, 0.05), lr=1e-6, betas=(0.9, 0.95), fused=True)
            for _ in range(2):
                topt.step()
            opt_torch_ms = timed(topt.step, 5) / 5
            del opt, topt
        except Exception as e:  # noqa
            if rank == 0:
                print(f"[bench] optimizer timing skipped ({type(e).__name__}: {e})", file=sys.stderr)

    # the north star's encoder target: forward + backward of forward_encoder alone (patch-embed, masking, 24 ViT-L blocks
    # at S = 410, final norm), as one CUDA graph, against the tensor-pipe peak
    enc = None                                                   # (after the optimizer timing: it drops the step's gradients)
    if world == 1 and not args.no_extras:
        try:
            enc = encoder_step(model, vol, pk, timed, args.steps)
        except Exception as e:  # noqa
            print(f"[bench] encoder-step timing skipped ({type(e).__name__}: {e})", file=sys.stderr)
    model = vol = host_vol = None
    sb.close()

    # BASELINE configs[2] (cfg-3): the same step on 60x256x256 volumes (S_dec = 5121, keep = 511), same N, same build
    cfg3 = None
    other = 60 if FRAMES == 48 else 48
    if not args.no_extras:
        try:
            sb3 = StepBench(other, args, dev, world, rank, lib)
            for _ in range(3):
                sb3.run()
            ms3 = timed(sb3.run, args.steps) / args.steps
            cfg3 = {"workload": workload_name(other), "value": world * BATCH / (ms3 / 1e3), "unit": "volumes/s",
                    "ms_per_step": ms3, "steps": args.steps, "cuda_graph": sb3.graph is not None,
                    "step_tflops_per_gpu": GF_PER_VOLUME[other] * BATCH / ms3,
                    "step_frac_of_bf16_sustained": GF_PER_VOLUME[other] * BATCH / ms3 / pk["bf16_sustained"]}
            sb3.close()
            del sb3
        except Exception as e:  # noqa
            if rank == 0:
                print(f"[bench] {other}-frame configuration skipped ({type(e).__name__}: {e})", file=sys.stderr)

    # the GPU incumbent on the same box: the reference's own flash model (flash_attn Blocks + nn.Conv3d under bf16 autocast,
    # oracle/gpu_incumbent.py — test infrastructure, never on the product path), same config, same precision
    incumbent = None
    if world == 1 and not args.no_extras and not args.no_incumbent:
        incumbent = incumbent_block(dev, timed, args.steps)

    if rank == 0:
        gf = GF_PER_VOLUME[FRAMES]
        line = {
            "metric": "3D-MAE pretrain volumes/sec", "value": value, "unit": "volumes/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": workload_name(FRAMES), "global_batch": world * BATCH, "parallelism": f"dp{world}",
                       "step": "forward + backward" + (" + NCCL gradient all-reduce overlapped with backward" if world > 1 else "")
                               + "; optimizer excluded (SURVEY §8d), see optimizer_ms",
                       "cuda_graph": bool(graph_used),
                       "l2": "per-step working set (1.3 GB fp32 weights + bf16 shadows + ~10 GB activations) >> 126 MB L2; no explicit flush"},
            "e2e": {"value": e2e_value, "unit": "volumes/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": BATCH * FRAMES * IMG * IMG * 4,
                    "d2h_bytes_per_step": 4,
                    "pipeline": "upload of step i+1 (pinned, copy stream, staging buffer) overlaps step i; one H2D per timed step"},
            "gpu_launches": launches_per_step * args.steps,
            "gpu_launches_per_step": launches_per_step,
            "clocks": clocks,
            "step_tflops_per_gpu": gf * BATCH / ms_step,
            "step_frac_of_bf16_sustained": gf * BATCH / ms_step / pk["bf16_sustained"],
            "roofline": roof,
            "encoder_step": enc,
            "cfg3" if other == 60 else "cfg2": cfg3,
            "comm": comm,
            "incumbent": incumbent,
            "optimizer_ms": opt_ms,
            "optimizer_torch_fused_ms": opt_torch_ms,
            "loss": last_loss,
            "peaks": pk["src"],
        }
        if world == 1 and not args.no_cpu_baseline:
            r = run_cpu(1, 0, budget_s=60.0)
            line["cpu_baseline"] = {"value": r["value"], "unit": "volumes/s", "cores": r["cores"], "kind": "port",
                                    "sample": r["sample"]}
        print(json.dumps(line), flush=True)
    sys.stdout.flush()
    sys.stderr.flush()
    # orderly exit (interpreter exit hooks run): every graph that held NCCL kernels has been destroyed above, so the
    # communicator can be torn down; a hung teardown is cut short by the exit watchdog rather than by skipping the hooks
    torch.cuda.synchronize()
    if world > 1:
        barrier()
        threading.Thread(target=_exit_watchdog, args=(60,), daemon=True).start()
        dist.destroy_process_group()


def _exit_watchdog(seconds):
    time.sleep(seconds)
    print(f"[bench] process-group teardown still running after {seconds} s; leaving", file=sys.stderr, flush=True)
    os._exit(0)


def incumbent_block(dev, timed, steps):
    """Same-box numbers of the stack the reference runs on a GPU (SURVEY §2.2 'bar'): its flash model under bf16 autocast,
    eager like the reference's loop and as a CUDA graph, plus flash_attn's kernels alone at the two attention shapes."""
    try:
        from oracle import gpu_incumbent as G
        return G.bench_block(FRAMES, IMG, BATCH, MASK, dev, timed, min(steps, 10))
    except Exception as e:  # noqa
        return {"unavailable": f"{type(e).__name__}: {e}"[:300]}


# algorithmic GF per volume of the encoder blocks, forward (24 x (24 S dim^2 + 4 S^2 dim), S = keep + 1)
ENC_GF_FWD = {48: 247.6 + 16.5, 60: 309.2 + 25.8}


def encoder_step(model, vol, pk, timed, steps):
    """forward_encoder + its backward (seeded with a fixed dlatent) on the bench batch, replayed as one CUDA graph."""
    model.zero_grad(set_to_none=True)
    with torch.no_grad():
        model._rt.shadows.begin_step()
        lat, _, _ = model.forward_encoder(vol, MASK)
    dlat = torch.randn(lat.shape, device=vol.device).to(lat.dtype) * 1e-3

    def enc_step(recast=True):
        model.zero_grad(set_to_none=True)
        if recast:
            model._rt.shadows.begin_step()                      # the fp32 -> bf16 weight casts are part of the captured step
        latent, _, _ = model.forward_encoder(vol, MASK)
        latent.backward(dlat)

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            enc_step()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        enc_step()
    for _ in range(3):
        graph.replay()
    ms = timed(graph.replay, steps) / steps
    # the same step when the optimizer has already emitted the bf16 shadows (optim.FusedAdamW(shadows=...): what a real
    # training loop of this package runs; the casts above are what autocast pays in the reference) — informational
    ms_nocast = None
    try:
        graph2 = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph2):
            enc_step(recast=False)
        for _ in range(3):
            graph2.replay()
        ms_nocast = timed(graph2.replay, steps) / steps
        del graph2
    except Exception as e:  # noqa
        print(f"[bench] encoder step without weight casts skipped ({type(e).__name__}: {e})", file=sys.stderr)
    model.zero_grad(set_to_none=True)
    gf = 3.0 * ENC_GF_FWD[FRAMES] * BATCH                       # bwd = 2 x fwd
    tf = gf / ms
    return {"what": f"forward_encoder fwd+bwd, batch {BATCH}, S = {lat.shape[1] + 1}, 24 blocks (patch-embed / masking / LN kernels "
                    "run inside the timed region; only the blocks' GEMM + attention FLOPs are counted)",
            "ms": ms, "volumes_per_s": BATCH / (ms / 1e3), "tflops": tf, "frac_of_bf16_sustained": tf / pk["bf16_sustained"],
            "frac_of_bf16_burst": tf / pk["bf16_burst"], "cuda_graph": True,
            "shadows_current": None if ms_nocast is None else {
                "what": "same step without the fp32 -> bf16 weight casts (bf16 shadows emitted by FusedAdamW, as in a training loop)",
                "ms": ms_nocast, "frac_of_bf16_sustained": gf / ms_nocast / pk["bf16_sustained"]}}


def dominant_kernel_roofline(ops, dev, pk):
    """Decoder attention backward (B=8, S=4097, 16 heads x 32): the largest single share of the step
    (profiles/r1_step_launches.md), timed alone with CUDA events on the launching stream.  `achieved` counts the
    ALGORITHMIC FLOPs of SURVEY §8d (bwd = 2 x 4 S^2 dim, no recompute); the same object also reports the kernel
    against its SFU (ex2) bound, which is the tighter roofline for head_dim 32."""
    from octcubem_b200._lib import EPI_BIAS, GEMM_NT, OCT_BF16
    B, S, H, d = BATCH, (FRAMES // 3) * 256 + 1, 16, 32
    qkv = (torch.randn(B, S, 3 * H * d, device=dev) * 0.5).bfloat16()
    dout = torch.randn(B, S, H * d, device=dev).bfloat16()
    out, lse = ops.attn_fwd(qkv, H, d, OCT_BF16)

    def timeit(fn, n=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(n):
            fn()
        e.record()
        torch.cuda.synchronize()
        return s.elapsed_time(e) / n

    ms = timeit(lambda: ops.attn_bwd(qkv, out, dout, lse, H, d, OCT_BF16))
    ms_fwd = timeit(lambda: ops.attn_fwd(qkv, H, d, OCT_BF16))
    flops = 2.0 * 4.0 * S * S * (H * d) * B
    achieved = flops / ms / 1e9
    n_exp = float(B) * H * S * S                       # one ex2 per score element per pass
    sfu_peak = 16.0 * 148 * 1.965e9                    # MUFU: 16 ex2 / clk / SM at the max SM clock
    # second data point: the biggest GEMM of the decoder block (fc1 + GELU epilogue), tensor-core bound
    M, N, K = B * S, 2048, 512
    a = torch.randn(M, K, device=dev).bfloat16(); w = torch.randn(N, K, device=dev).bfloat16()
    bias = torch.zeros(N, device=dev); o = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
    ms_gemm = timeit(lambda: ops.gemm(GEMM_NT, a, w, M, N, K, torch.bfloat16, EPI_BIAS, bias=bias, out=o, compute=OCT_BF16))
    return {"kernel": "attn_bwd_tc_kernel<32> (+delta, dq-convert), decoder shape B8 S4097 H16 d32", "bound": "tensor",
            "achieved": achieved, "peak": pk["bf16_burst"], "unit": "TFLOP/s", "frac": achieved / pk["bf16_burst"],
            "traffic": 305.6e6, "traffic_note": "dram read+write per launch from ncu --set full of the final kernel (profiles/r2_attention_ncu.md: 208.3 MB read + 97.3 MB written)",
            "ms_per_launch": ms, "peak_source": pk["src"],
            "sfu_bound": {"ex2_per_launch": n_exp, "achieved_gex2_s": n_exp / ms / 1e6, "peak_gex2_s": sfu_peak / 1e9,
                          "frac": n_exp / (ms * 1e-3) / sfu_peak,
                          "note": "head_dim 32: 128 MMA flop per score element, so ex2 throughput (16/clk/SM), not the tensor pipe, bounds this kernel"},
            "attn_fwd": {"ms_per_launch": ms_fwd, "tflops": 4.0 * S * S * (H * d) * B / ms_fwd / 1e9,
                         "sfu_frac": n_exp / (ms_fwd * 1e-3) / sfu_peak},
            "gemm_fc1": {"shape": [M, N, K], "ms_per_launch": ms_gemm, "tflops": 2.0 * M * N * K / ms_gemm / 1e9,
                         "frac_of_bf16_burst": 2.0 * M * N * K / ms_gemm / 1e9 / pk["bf16_burst"]},
            "note": "algorithmic FLOPs (no recompute counted); inputs (100 MB qkv + 34 MB dout + 34 MB out) exceed the 126 MB L2"}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="headline + e2e + roofline only (no encoder / cfg-3 / optimizer / incumbent legs)")
    ap.add_argument("--no-incumbent", action="store_true")
    ap.add_argument("--frames", type=int, default=FRAMES, choices=sorted(GF_PER_VOLUME),
                    help="48 = BASELINE.json configs[1] (the default and the driver's workload); 60 = the cfg-3 volume size")
    a = ap.parse_args()
    FRAMES = a.frames
    if a.impl == "reference":
        main_reference(a)
    else:
        main_b200(a)
