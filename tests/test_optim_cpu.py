"""CPU checks of the optimizer-side host logic (octcubem_b200/optim.py): the reference's weight-decay grouping
(custom_util/misc.py:678-696), its half-cycle cosine schedule (custom_util/lr_sched.py:10-28), and the fail-loud rule —
FusedAdamW has no CPU path."""
import math

import pytest
import torch

from octcubem_b200 import optim


class _Net(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.cls_token = torch.nn.Parameter(torch.zeros(1, 1, 8))
        self.fc = torch.nn.Linear(8, 4)
        self.norm = torch.nn.LayerNorm(4)
        self.frozen = torch.nn.Parameter(torch.zeros(3), requires_grad=False)


def test_add_weight_decay_groups_like_the_reference():
    net = _Net()
    no_decay, decay = optim.add_weight_decay(net, 0.05)
    names = {id(p): k for k, p in net.named_parameters()}
    assert no_decay["weight_decay"] == 0.0 and decay["weight_decay"] == 0.05
    assert sorted(names[id(p)] for p in no_decay["params"]) == ["fc.bias", "norm.bias", "norm.weight"]
    assert sorted(names[id(p)] for p in decay["params"]) == ["cls_token", "fc.weight"]
    # skip_list and bias_wd: 1-D tensors decay when bias_wd is set, but `*.bias` never does; frozen tensors are dropped
    no_decay, decay = optim.add_weight_decay(net, 0.05, skip_list=("cls_token",), bias_wd=True)
    assert sorted(names[id(p)] for p in no_decay["params"]) == ["cls_token", "fc.bias", "norm.bias"]
    assert sorted(names[id(p)] for p in decay["params"]) == ["fc.weight", "norm.weight"]


def test_cosine_schedule_matches_the_reference_formula():
    opt = torch.optim.SGD([{"params": [torch.nn.Parameter(torch.zeros(1))]},
                           {"params": [torch.nn.Parameter(torch.zeros(1))], "lr_scale": 0.5}], lr=1.0)
    lr, min_lr, warm, epochs = 1.6e-3, 1e-6, 5, 100
    for epoch in (0.0, 2.5, 5.0, 17.25, 99.9):
        got = optim.adjust_learning_rate(opt, epoch, lr, min_lr, warm, epochs)
        want = lr * epoch / warm if epoch < warm else min_lr + (lr - min_lr) * 0.5 * (1 + math.cos(math.pi * (epoch - warm) / (epochs - warm)))
        assert got == pytest.approx(want, rel=1e-12)
        assert opt.param_groups[0]["lr"] == got and opt.param_groups[1]["lr"] == got * 0.5


def test_fused_adamw_refuses_cpu_parameters():
    p = torch.nn.Parameter(torch.randn(4))
    p.grad = torch.randn(4)
    with pytest.raises(RuntimeError):
        optim.FusedAdamW([p]).step()


def test_cosine_schedule_object_equals_adjust_learning_rate():
    """CosineSchedule.lr_at_step(k) = the reference rule at the first iteration of accumulation group k (engine_pretrain.py:87-91)."""
    opt = torch.optim.SGD([torch.nn.Parameter(torch.zeros(1))], lr=1.0)
    accum, iters_per_epoch = 4, 250
    s = optim.CosineSchedule(lr=1.6e-3, min_lr=1e-6, warmup_epochs=5, epochs=100, epochs_per_step=accum / iters_per_epoch)
    for k in (1, 2, 63, 313, 314, 5000, 6250):
        data_iter_step = (k - 1) * accum
        want = optim.adjust_learning_rate(opt, data_iter_step / iters_per_epoch, 1.6e-3, 1e-6, 5, 100)
        assert s.lr_at_step(k) == pytest.approx(want, rel=1e-9, abs=1e-15)


def test_mask_ratio_and_K_schedulers():
    """main_pretrain_oph_joint_2d512_flash_attn.py:53-67: flat during the warm-up, then linear."""
    from octcubem_b200.engine_pretrain import K_scheduler, mask_ratio_2d_scheduler
    assert mask_ratio_2d_scheduler(0) == 0.75 and mask_ratio_2d_scheduler(10) == 0.75
    assert mask_ratio_2d_scheduler(55) == pytest.approx(0.75 + 45 * 0.10 / 90)
    assert mask_ratio_2d_scheduler(100) == pytest.approx(0.85)
    assert mask_ratio_2d_scheduler(30, epoch_offset=20, all_epoch=120) == 0.75
    assert K_scheduler(5) == 0.7 and K_scheduler(100) == pytest.approx(0.3) and K_scheduler(55) == pytest.approx(0.5)
