"""GPU: the graph-replayed training step (octcubem_b200/engine_pretrain.JointPretrainStep, mirror of the loop body of
Pre-training/engine_pretrain.py:83-161) against a plain eager loop (autograd + torch.optim.AdamW + clip_grad_norm_ + the
host-side cosine schedule of lr_sched.py) on the same module, inputs and masks."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from octcubem_b200 import models_mae, optim  # noqa: E402
from octcubem_b200.engine_pretrain import JointPretrainStep  # noqa: E402
from oracle import mae3d_oracle as O  # noqa: E402
from oracle.gen_golden import TOY, toy_inputs  # noqa: E402

DEV = "cuda:0"


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def build(sd, precision):
    m = models_mae.MaskedAutoencoderViT(**TOY.ref_kwargs(), use_flash_attn=True, precision=precision,
                                        norm_layer=lambda d: torch.nn.LayerNorm(d, eps=1e-6)).to(DEV)
    m.load_state_dict(sd, strict=True)
    return m


@pytest.mark.parametrize("joint", [True, False])
def test_graph_replayed_steps_match_an_eager_torch_loop(joint):
    sd, _, _ = toy_inputs()
    sched = dict(lr=3e-3, min_lr=1e-5, warmup_epochs=1.0, epochs=3.0)   # lr per step: 0, 1.5e-3, 3e-3, 2.6e-3, 1.5e-3, 4.5e-4
    eng_model, ref_model = build(sd, "fp32"), build(sd, "fp32")
    opt = optim.FusedAdamW(optim.add_weight_decay(eng_model, 0.05), betas=(0.9, 0.95),
                           schedule=optim.CosineSchedule(**sched, epochs_per_step=0.5))
    engine = JointPretrainStep(eng_model, opt, mask_ratio=0.9, clip_grad=0.5, use_graph=True, warm_steps=2)
    ropt = torch.optim.AdamW(optim.add_weight_decay(ref_model, 0.05), lr=1.0, betas=(0.9, 0.95))
    try:
        for k in range(1, 7):                                   # steps 1-2 eager, 3 captures + replays, 4-6 replay
            vol = O.synthetic_volume(2, 12, 64, 64, seed=10 + k, zero_pad_frames=1).to(DEV)
            noise = O.synthetic_noise(2, 64, seed=20 + k).to(DEV)
            img = O.synthetic_volume(2, 3, 128, 128, seed=30 + k, zero_pad_frames=0).to(DEV) if joint else None
            noise2 = O.synthetic_noise(2, 64, seed=40 + k).to(DEV) if joint else None
            res = engine(vol.view(1, 2, 1, 12, 64, 64), img, mask_ratio_2d=0.75, noise=noise, noise_2d=noise2)
            # the eager reference loop (engine_pretrain.py:87-173 with torch's optimizer)
            optim.adjust_learning_rate(ropt, (k - 1) * 0.5, sched["lr"], sched["min_lr"], sched["warmup_epochs"], sched["epochs"])
            ref_model.zero_grad(set_to_none=True)
            (loss, fl), _, _ = ref_model(vol, mask_ratio=0.9, frame_loss=True, noise=noise)
            total = loss
            if joint:
                loss2, _, _ = ref_model(img, mask_ratio=0.75, noise=noise2)
                total = loss + loss2
            total.backward()
            norm = torch.nn.utils.clip_grad_norm_(ref_model.parameters(), 0.5)
            ropt.step()
            got = res.check_finite()
            # Adam divides by sqrt(v): round-off level differences between the two gradient paths (in-place sinks with atomics
            # vs autograd sums) are amplified for near-zero entries, so parameters agree to ~1e-4, not to fp32 round-off; a
            # learning rate or bias correction frozen at capture time (step 3) would be off by 10 % and more from step 4 on
            assert got["loss"] == pytest.approx(float(loss.detach()), rel=2e-4), k
            assert got["loss_all"] == pytest.approx(float(total), rel=2e-4), k
            assert got["grad_norm"] == pytest.approx(float(norm), rel=1e-3), k
            assert rel(res.frame_loss, fl.detach()) < 2e-4
            if joint:
                assert got["loss_2d"] == pytest.approx(float(loss2), rel=2e-4), k
            for (name, p), (_, r) in zip(eng_model.named_parameters(), ref_model.named_parameters()):
                assert rel(p.detach(), r.detach()) < 1e-3, (k, name)
        step, dev_lr = opt.clock_state()
        assert step == 6 and dev_lr == pytest.approx(opt.schedule.lr_at_step(6), rel=1e-5)
        ent = next(iter(engine._entries.values()))
        assert ent["graph"] is not None and ent["calls"] == 2
    finally:
        engine.close()


def test_graph_replayed_steps_match_the_oracle_loop():
    """The engine (forward, backward, clip, AdamW on the device clock; eager warm steps, capture, replays) against the loop
    body of the reference's train_one_epoch (engine_pretrain.py:87-173) run on the CPU ORACLE: oracle forward + autograd
    (pinned to the unmodified reference in tests/test_oracle.py), torch's clip_grad_norm_ and AdamW on CPU parameters."""
    sd, _, _ = toy_inputs()
    sched = dict(lr=3e-3, min_lr=1e-5, warmup_epochs=1.0, epochs=3.0)
    eng_model = build(sd, "fp32")
    opt = optim.FusedAdamW(optim.add_weight_decay(eng_model, 0.05), betas=(0.9, 0.95),
                           schedule=optim.CosineSchedule(**sched, epochs_per_step=0.5))
    engine = JointPretrainStep(eng_model, opt, mask_ratio=0.9, clip_grad=0.5, use_graph=True, warm_steps=2)
    # CPU parameter container with the reference's names (never run: the product model has no CPU path)
    cpu = models_mae.MaskedAutoencoderViT(**TOY.ref_kwargs(), use_flash_attn=True, precision="fp32",
                                          norm_layer=lambda d: torch.nn.LayerNorm(d, eps=1e-6))
    cpu.load_state_dict(sd, strict=True)
    ropt = torch.optim.AdamW(optim.add_weight_decay(cpu, 0.05), lr=1.0, betas=(0.9, 0.95))
    try:
        for k in range(1, 6):
            vol = O.synthetic_volume(2, 12, 64, 64, seed=110 + k, zero_pad_frames=1)
            noise = O.synthetic_noise(2, 64, seed=120 + k)
            res = engine(vol.to(DEV).view(1, 2, 1, 12, 64, 64), None, noise=noise.to(DEV))
            optim.adjust_learning_rate(ropt, (k - 1) * 0.5, sched["lr"], sched["min_lr"], sched["warmup_epochs"], sched["epochs"])
            cur = {n: p.detach() for n, p in cpu.state_dict().items()}
            (out, grads) = O.forward_backward(TOY, cur, vol, 0.9, noise, frame_loss=True)
            (loss, fl), _, _ = out
            for n, p in cpu.named_parameters():
                p.grad = grads.get(n)                              # the high-res patch embedding takes no part in a 3D step
            norm = torch.nn.utils.clip_grad_norm_([p for p in cpu.parameters() if p.grad is not None], 0.5)
            ropt.step()
            got = res.check_finite()
            assert got["loss"] == pytest.approx(float(loss.detach()), rel=2e-4), k
            assert got["grad_norm"] == pytest.approx(float(norm), rel=1e-3), k
            assert rel(res.frame_loss, fl.detach()) < 2e-4
            for (name, p), (_, r) in zip(eng_model.named_parameters(), cpu.named_parameters()):
                assert rel(p.detach(), r.detach()) < 1e-3, (k, name)
    finally:
        engine.close()


def test_engine_uint8_cubes_equal_the_fp32_path():
    """A uint8 cube through ops.ingest_u8 (eager steps: fresh tensor; replayed steps: straight into the graph's static input)
    gives the same losses as feeding the CPU pipeline's fp32 volume."""
    sd, _, _ = toy_inputs()
    models = [build(sd, "fp32"), build(sd, "fp32")]
    engines = []
    for m in models:
        opt = optim.FusedAdamW(optim.add_weight_decay(m, 0.05), betas=(0.9, 0.95),
                               schedule=optim.CosineSchedule(1e-3, 1e-5, 1.0, 3.0, 0.5))
        engines.append(JointPretrainStep(m, opt, mask_ratio=0.9, clip_grad=1.0, use_graph=True, warm_steps=2))
    try:
        for k in range(1, 5):
            g = torch.Generator().manual_seed(50 + k)
            cube = torch.randint(0, 256, (2, 10, 64, 64), generator=g, dtype=torch.uint8)      # 10 frames -> padded to 12
            flip_t = torch.tensor([k % 2, 1], dtype=torch.uint8)
            noise = O.synthetic_noise(2, 64, seed=60 + k).to(DEV)
            vol = torch.cat([torch.zeros(2, 1, 64, 64), cube.float().div(255), torch.zeros(2, 1, 64, 64)], 1)
            vol = torch.stack([v.flip(0) if f else v for v, f in zip(vol, flip_t)]).unsqueeze(1)
            a = engines[0](cube.to(DEV), noise=noise, flips=(flip_t.to(DEV), None)).check_finite()
            b = engines[1](vol.to(DEV), noise=noise).check_finite()
            assert a["loss"] == pytest.approx(b["loss"], rel=1e-6) and a["grad_norm"] == pytest.approx(b["grad_norm"], rel=1e-5), k
    finally:
        for e in engines:
            e.close()


def test_engine_refuses_a_host_side_schedule_under_graphs():
    sd, _, _ = toy_inputs()
    m = build(sd, "fp32")
    with pytest.raises(ValueError, match="schedule"):
        JointPretrainStep(m, optim.FusedAdamW(optim.add_weight_decay(m, 0.05)), use_graph=True)
    with pytest.raises(ValueError):
        JointPretrainStep(m, optim.FusedAdamW(optim.add_weight_decay(m, 0.05)), use_graph=False, accum_iter=0)


@pytest.mark.parametrize("accum", [2, 3])
def test_gradient_accumulation_matches_the_reference_loop(accum):
    """accum_iter > 1 (engine_pretrain.py:163-173): `loss /= accum_iter`, backward every call, optimizer + zero_grad on every
    accum_iter-th call, the learning rate set at the first call of each group — against the eager torch loop.  Groups 1-2 run
    eagerly (bucket discovery inside an accumulation group, then the in-place sinks accumulating), groups 3-4 replay one
    captured graph per micro-step phase."""
    sd, _, _ = toy_inputs()
    sched = dict(lr=3e-3, min_lr=1e-5, warmup_epochs=1.0, epochs=3.0)
    eng_model, ref_model = build(sd, "fp32"), build(sd, "fp32")
    opt = optim.FusedAdamW(optim.add_weight_decay(eng_model, 0.05), betas=(0.9, 0.95),
                           schedule=optim.CosineSchedule(**sched, epochs_per_step=0.5))
    engine = JointPretrainStep(eng_model, opt, mask_ratio=0.9, clip_grad=0.5, use_graph=True, warm_steps=2, accum_iter=accum)
    ropt = torch.optim.AdamW(optim.add_weight_decay(ref_model, 0.05), lr=1.0, betas=(0.9, 0.95))
    try:
        it = 0
        for group in range(1, 5):
            optim.adjust_learning_rate(ropt, (group - 1) * 0.5, sched["lr"], sched["min_lr"], sched["warmup_epochs"], sched["epochs"])
            for micro in range(accum):
                it += 1
                vol = O.synthetic_volume(2, 12, 64, 64, seed=10 + it, zero_pad_frames=1).to(DEV)
                noise = O.synthetic_noise(2, 64, seed=20 + it).to(DEV)
                res = engine(vol, noise=noise)
                (loss, fl), _, _ = ref_model(vol, mask_ratio=0.9, frame_loss=True, noise=noise)
                (loss / accum).backward()
                assert float(res.loss) == pytest.approx(float(loss.detach()), rel=2e-4), (group, micro)
            norm = torch.nn.utils.clip_grad_norm_(ref_model.parameters(), 0.5)
            ropt.step()
            ropt.zero_grad(set_to_none=True)
            assert float(res.grad_norm) == pytest.approx(float(norm), rel=1e-3), group
            for (name, p), (_, r) in zip(eng_model.named_parameters(), ref_model.named_parameters()):
                assert rel(p.detach(), r.detach()) < 1e-3, (group, name)
        assert opt.clock_state()[0] == 4                                     # one optimizer step per group
        assert all(e["graph"] is not None for e in engine._entries.values())
    finally:
        engine.close()
