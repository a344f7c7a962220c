import math

import pytest
import torch

from deepspeed_b200.runtime.fp16.loss_scaler import CreateLossScaler, DynamicLossScaler
from deepspeed_b200.runtime.lr_schedules import (LRRangeTest, OneCycle, WarmupCosineLR, WarmupDecayLR, WarmupLR)


def _opt(lr=0.1):
    return torch.optim.SGD([torch.nn.Parameter(torch.zeros(1))], lr=lr, momentum=0.9)


def test_warmup_lr_linear_then_hold():
    o = _opt()
    s = WarmupLR(o, warmup_min_lr=0.0, warmup_max_lr=1.0, warmup_num_steps=10, warmup_type="linear")
    lrs = []
    for _ in range(15):
        s.step()
        lrs.append(o.param_groups[0]["lr"])
    assert lrs[0] == 0.0 and abs(lrs[5] - 0.5) < 1e-9 and all(abs(x - 1.0) < 1e-9 for x in lrs[10:])


def test_warmup_log_matches_formula():
    o = _opt()
    s = WarmupLR(o, warmup_min_lr=0.0, warmup_max_lr=2.0, warmup_num_steps=100)
    for _ in range(10):
        s.step()
    assert abs(o.param_groups[0]["lr"] - 2.0 * math.log(10) / math.log(100)) < 1e-9


def test_warmup_decay_reaches_zero():
    o = _opt()
    s = WarmupDecayLR(o, total_num_steps=20, warmup_max_lr=1.0, warmup_num_steps=10, warmup_type="linear")
    for _ in range(21):
        s.step()
    assert o.param_groups[0]["lr"] == 0.0


def test_cosine_endpoints():
    o = _opt(lr=1.0)
    s = WarmupCosineLR(o, total_num_steps=110, warmup_num_steps=10, cos_min_ratio=0.1, warmup_type="linear")
    vals = []
    for _ in range(111):
        s.step()
        vals.append(o.param_groups[0]["lr"])
    assert abs(max(vals) - 1.0) < 1e-2 and abs(vals[-1] - 0.1) < 1e-6


def test_lr_range_test_staircase_and_continuous():
    o = _opt()
    s = LRRangeTest(o, lr_range_test_min_lr=0.01, lr_range_test_step_size=5, lr_range_test_step_rate=1.0,
                    lr_range_test_staircase=True)
    got = []
    for _ in range(10):
        s.step()
        got.append(round(o.param_groups[0]["lr"], 6))
    assert got[:4] == [0.01] * 4 and got[4] == 0.02


def test_one_cycle_up_down_and_momentum():
    o = _opt()
    s = OneCycle(o, cycle_min_lr=0.1, cycle_max_lr=1.0, cycle_first_step_size=10, cycle_min_mom=0.8, cycle_max_mom=0.9)
    lrs, moms = [], []
    for _ in range(20):
        s.step()
        lrs.append(o.param_groups[0]["lr"])
        moms.append(o.param_groups[0]["momentum"])
    assert abs(max(lrs) - 1.0) < 1e-9 and lrs.index(max(lrs)) == 9
    assert abs(min(moms) - 0.8) < 1e-9 and moms.index(min(moms)) == 9
    sd = s.state_dict()
    s2 = OneCycle(_opt(), cycle_min_lr=0.1, cycle_max_lr=1.0, cycle_first_step_size=10)
    s2.load_state_dict(sd)
    assert s2.last_batch_iteration == s.last_batch_iteration


def test_dynamic_loss_scaler_hysteresis_and_window():
    s = DynamicLossScaler(init_scale=2**8, scale_window=3, delayed_shift=2, min_scale=1)
    s.update_scale(True)  # first overflow consumes hysteresis
    assert s.cur_scale == 2**8
    s.update_scale(True)
    assert s.cur_scale == 2**7
    for _ in range(3):
        s.update_scale(False)
    assert s.cur_scale == 2**8
    s2 = DynamicLossScaler(init_scale=2, min_scale=1, delayed_shift=1, raise_error_at_min_scale=True)
    s2.update_scale(True)
    with pytest.raises(Exception):
        s2.update_scale(True)


def test_create_loss_scaler_static_for_bf16():
    s = CreateLossScaler(torch.bfloat16, static_loss_scale=128, dynamic_scaling=True, dynamic_loss_args=None)
    assert s.cur_scale == 1.0 and not s.dynamic
