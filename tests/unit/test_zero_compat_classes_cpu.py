"""The reference's class names / constructors (DeepSpeedZeroOptimizer, DeepSpeedZeroOptimizer_Stage3,
PartitionedParameterCoordinator) drive the unified sharded optimizer and match plain AdamW."""
import copy

import torch
from torch import nn

from tests.common import run_distributed


def _mlp():
    torch.manual_seed(0)
    return nn.Sequential(nn.Linear(8, 16), nn.Tanh(), nn.Linear(16, 4))


def _stage12(partition_grads):
    import torch.distributed as td
    from deepspeed_b200.runtime.base_optimizer import ZeROOptimizer
    from deepspeed_b200.runtime.zero.stage_1_and_2 import DeepSpeedZeroOptimizer
    r, w = td.get_rank(), td.get_world_size()
    model, ref = _mlp(), _mlp()
    ropt = torch.optim.AdamW(ref.parameters(), lr=1e-2, weight_decay=0.0)
    opt = DeepSpeedZeroOptimizer(torch.optim.AdamW(model.parameters(), lr=1e-2, weight_decay=0.0), partition_grads=partition_grads,
                                 clip_grad=0.0, device="cpu")
    assert isinstance(opt, ZeROOptimizer) and opt.stage == (2 if partition_grads else 1)
    torch.manual_seed(1)
    data = torch.randn(4 * w, 8)
    for _ in range(3):
        x = data[r * 4:(r + 1) * 4]
        opt.backward(model(x).pow(2).mean())
        opt.step()
        ropt.zero_grad()
        ref(data).pow(2).mean().backward()
        ropt.step()
    for a, b in zip(model.parameters(), ref.parameters()):
        assert torch.allclose(a, b, atol=1e-5), (a - b).abs().max()


def test_stage1_and_stage2_class():
    run_distributed(_stage12, 2, (False, ))
    run_distributed(_stage12, 2, (True, ))


def _stage3():
    import torch.distributed as td
    from deepspeed_b200.runtime.zero.partitioned_param_coordinator import PartitionedParameterCoordinator, ZeRoTraceMode
    from deepspeed_b200.runtime.zero.stage3 import DeepSpeedZeroOptimizer_Stage3
    r, w = td.get_rank(), td.get_world_size()
    model, ref = _mlp(), _mlp()
    ropt = torch.optim.AdamW(ref.parameters(), lr=1e-2, weight_decay=0.0)
    opt = DeepSpeedZeroOptimizer_Stage3(model, torch.optim.AdamW(model.parameters(), lr=1e-2, weight_decay=0.0),
                                        param_persistence_threshold=0, device="cpu")
    coord = PartitionedParameterCoordinator(opt)
    assert coord.trace_mode is ZeRoTraceMode.RECORD
    lin0 = model[0]
    assert lin0.weight.numel() == 0  # partitioned away
    coord.fetch_sub_module(lin0)
    assert lin0.weight.shape == (16, 8) and coord.available_parameter_numel > 0
    coord.release_sub_module(lin0)
    assert lin0.weight.numel() == 0
    opt.reset_step()
    torch.manual_seed(1)
    data = torch.randn(4 * w, 8)
    for _ in range(3):
        x = data[r * 4:(r + 1) * 4]
        opt.backward(model(x).pow(2).mean())
        opt.step()
        coord.reset_step()
        ropt.zero_grad()
        ref(data).pow(2).mean().backward()
        ropt.step()
    assert coord.is_complete_trace()
    assert len(coord.construct_parameter_trace_from_module_trace()) == 4
    coord.release_and_reset_all(model)
    for p, q in zip(model.parameters(), ref.parameters()):
        full = opt.get_full_hp_param(p)
        assert torch.allclose(full.view_as(q), q, atol=1e-5)


def test_stage3_class_and_coordinator():
    run_distributed(_stage3, 2)
