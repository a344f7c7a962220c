"""Regression tests for the round-1 advisor findings: loss-scale handling on the fused path, a unit reduced
twice inside one backward, changing GAS after construction, module-only ZeRO-3 loads, replicated optimizer shards."""
import copy
import os

import pytest
import torch
from torch import nn

from tests.common import run_distributed
from tests.unit.simple_model import SimpleModel, base_config, make_batch


def _static_scale_worker():
    import deepspeed_b200 as ds
    from deepspeed_b200.utils import safe_get_full_fp32_param
    deltas = []
    for scale in (1.0, 128.0):
        torch.manual_seed(0)
        model = SimpleModel()
        before = copy.deepcopy(model)
        cfg = base_config(2, "fp16", opt="SGD", lr=0.1)
        cfg["fp16"] = {"enabled": True, "loss_scale": scale}
        eng, *_ = ds.initialize(model=model, config=cfg)
        assert not eng.optimizer.fused_in_backward, "fp16 must never take the fused-in-backward path"
        g = torch.Generator().manual_seed(1)
        x, y = make_batch(1, 4, g)
        loss = eng(x.half(), y)
        eng.backward(loss)
        eng.step()
        d = max((safe_get_full_fp32_param(p).cpu() - q.float()).abs().max().item()
                for p, q in zip(model.parameters(), before.parameters()))
        deltas.append(d)
    # a static loss scale must not change the size of the update (it is divided out before the step)
    assert abs(deltas[0] - deltas[1]) < 0.05 * deltas[0] + 1e-3, deltas


def test_static_fp16_loss_scale_is_unscaled():
    run_distributed(_static_scale_worker, 1)


class _Twice(nn.Module):
    """One weight used inside a re-entrant checkpoint region AND outside it: its unit reduces twice per backward."""

    def __init__(self, use_ckpt):
        super().__init__()
        self.layers = nn.ModuleList([nn.Linear(8, 8), nn.Linear(8, 8)])
        self.use_ckpt = use_ckpt

    def forward(self, x):
        from torch.utils.checkpoint import checkpoint
        l0 = self.layers[0]
        h = checkpoint(lambda t: torch.tanh(l0(t)), x, use_reentrant=True) if self.use_ckpt else torch.tanh(l0(x))
        h = self.layers[1](h)
        return (l0(h)**2).mean()


def _twice_worker(stage):
    import deepspeed_b200 as ds
    from deepspeed_b200.utils import safe_get_full_fp32_param
    out = []
    for use_ckpt in (False, True):
        torch.manual_seed(0)
        m = _Twice(use_ckpt)
        cfg = base_config(stage, opt="SGD", lr=0.1, clip=1.0)
        eng, *_ = ds.initialize(model=m, config=cfg)
        x = torch.randn(4, 8, generator=torch.Generator().manual_seed(3), requires_grad=True)
        loss = eng(x)
        eng.backward(loss)
        eng.step()
        out.append([safe_get_full_fp32_param(p).cpu().clone() for p in m.parameters()])
    for a, b in zip(*out):
        torch.testing.assert_close(a, b, atol=1e-6, rtol=1e-5)


@pytest.mark.parametrize("stage", [0, 2])
def test_unit_reduced_twice_accumulates(stage):
    run_distributed(_twice_worker, 1, (stage, ))


def _twice_fused_raises():
    import deepspeed_b200 as ds
    torch.manual_seed(0)
    m = _Twice(True)
    eng, *_ = ds.initialize(model=m, config=base_config(0, opt="Adam"))
    if not eng.optimizer.fused_in_backward:
        return  # host tier without a fused optimizer: nothing to check
    x = torch.randn(4, 8, requires_grad=True)
    with pytest.raises(RuntimeError, match="twice"):
        eng.backward(eng(x))


def test_fused_step_refuses_double_reduce():
    run_distributed(_twice_fused_raises, 1)


def _set_batch_worker():
    import deepspeed_b200 as ds
    from deepspeed_b200.utils import safe_get_full_fp32_param
    torch.manual_seed(0)
    model = SimpleModel()
    ref = copy.deepcopy(model)
    eng, *_ = ds.initialize(model=model, config=base_config(2))
    eng.set_train_batch_size(8)  # micro 4 x dp 1 -> GAS 2
    assert eng.gradient_accumulation_steps() == 2 and eng.optimizer.gas == 2
    assert not eng.optimizer.fused_in_backward and eng.optimizer.grad_arena is not None
    ropt = torch.optim.AdamW(ref.parameters(), lr=1e-2, weight_decay=0.01)
    g = torch.Generator().manual_seed(1)
    for it in range(4):
        x, y = make_batch(1, 4, g)
        eng.backward(eng(x, y))
        eng.step()
        (ref(x, y) / 2).backward()
        if it % 2 == 1:
            ropt.step()
            ropt.zero_grad()
    assert eng.global_steps == 2
    for p, q in zip(model.parameters(), ref.parameters()):
        torch.testing.assert_close(safe_get_full_fp32_param(p).cpu(), q.detach(), atol=1e-5, rtol=1e-4)


def test_set_train_batch_size_enables_accumulation():
    run_distributed(_set_batch_worker, 1)


def _forced_boundary_worker():
    import deepspeed_b200 as ds
    torch.manual_seed(0)
    eng, *_ = ds.initialize(model=SimpleModel(), config=base_config(1, gas=4))
    g = torch.Generator().manual_seed(1)
    x, y = make_batch(1, 4, g)
    eng.set_gradient_accumulation_boundary(True)  # step after ONE micro batch although GAS is 4
    assert eng.optimizer.is_gradient_accumulation_boundary()
    eng.backward(eng(x, y))
    eng.step()
    assert eng.global_steps == 1
    eng.set_gradient_accumulation_boundary(None)
    assert not eng.optimizer.is_gradient_accumulation_boundary()


def test_forced_gradient_accumulation_boundary_is_honoured():
    run_distributed(_forced_boundary_worker, 1)


def _module_only_worker(d, phase):
    import deepspeed_b200 as ds
    from deepspeed_b200.utils import safe_get_full_fp32_param
    torch.manual_seed(0 if phase == "save" else 99)
    eng, *_ = ds.initialize(model=SimpleModel(), config=base_config(3))
    if phase == "save":
        g = torch.Generator().manual_seed(1)
        x, y = make_batch(ds.comm.get_world_size(), 4, g)
        r = ds.comm.get_rank()
        eng.backward(eng(x[r * 4:(r + 1) * 4], y[r * 4:(r + 1) * 4]))
        eng.step()
        eng.save_checkpoint(d, tag="t")
        full = {n: safe_get_full_fp32_param(p).cpu() for n, p in eng.module.named_parameters()}
        if r == 0:
            torch.save(full, os.path.join(d, "expect.pt"))
    else:
        m_before = eng.optimizer.flat_opt.state_tensors()
        m_before = {k: v.clone() for k, v in m_before.items()}
        path, _ = eng.load_checkpoint(d, tag="t", load_module_only=True)
        assert path is not None
        exp = torch.load(os.path.join(d, "expect.pt"))
        for n, p in eng.module.named_parameters():
            torch.testing.assert_close(safe_get_full_fp32_param(p).cpu(), exp[n], atol=1e-6, rtol=1e-5)
        for k, v in eng.optimizer.flat_opt.state_tensors().items():  # optimizer moments untouched
            torch.testing.assert_close(v, m_before[k])


def test_zero3_load_module_only_restores_weights(tmp_path):
    d = str(tmp_path)
    run_distributed(_module_only_worker, 2, (d, "save"))
    run_distributed(_module_only_worker, 2, (d, "load"))


def _stage0_files_worker(d):
    import deepspeed_b200 as ds
    torch.manual_seed(0)
    eng, *_ = ds.initialize(model=SimpleModel(), config=base_config(0))
    g = torch.Generator().manual_seed(1)
    x, y = make_batch(2, 4, g)
    r = ds.comm.get_rank()
    eng.backward(eng(x[r * 4:(r + 1) * 4], y[r * 4:(r + 1) * 4]))
    eng.step()
    eng.save_checkpoint(d, tag="t")
    ds.comm.barrier()
    files = sorted(f for f in os.listdir(os.path.join(d, "t")) if f.endswith("_optim_states.pt"))
    assert len(files) == 1, files  # replicated optimizer: one shard, written by dp rank 0
    path, _ = eng.load_checkpoint(d, tag="t")
    assert path is not None


def test_stage0_writes_one_optimizer_shard(tmp_path):
    run_distributed(_stage0_files_worker, 2, (str(tmp_path), ))


def _twin_flow_worker():
    """offload_optimizer.ratio < 1: host + device optimizer domains must track a plain AdamW run (incl. clipping)."""
    import deepspeed_b200 as ds
    from deepspeed_b200.runtime.zero.multi import ZeroOptimizerGroup
    from deepspeed_b200.utils import safe_get_full_fp32_param
    torch.manual_seed(0)
    model = SimpleModel()
    ref = copy.deepcopy(model)
    cfg = base_config(3, clip=1.0)
    cfg["zero_optimization"]["offload_optimizer"] = {"device": "cpu", "ratio": 0.5}
    eng, opt, *_ = ds.initialize(model=model, config=cfg)
    assert isinstance(opt, ZeroOptimizerGroup) and [p.name for p in opt.parts] == ["twinflow:host", "twinflow:device"]
    assert opt.parts[0].offload_optimizer and not opt.parts[1].offload_optimizer
    n0 = sum(s.numel for u in opt.parts[0].units for s in u.slots)
    n1 = sum(s.numel for u in opt.parts[1].units for s in u.slots)
    assert 0.3 < n0 / (n0 + n1) < 0.8, (n0, n1)
    r, w = ds.comm.get_rank(), ds.comm.get_world_size()
    ropt = torch.optim.AdamW(ref.parameters(), lr=1e-2, weight_decay=0.01)
    g = torch.Generator().manual_seed(1)
    for _ in range(4):
        x, y = make_batch(w, 4, g)
        eng.backward(eng(x[r * 4:(r + 1) * 4], y[r * 4:(r + 1) * 4]))
        eng.step()
        ref(x, y).backward()
        torch.nn.utils.clip_grad_norm_(ref.parameters(), 1.0)
        ropt.step()
        ropt.zero_grad()
    for p, q in zip(model.parameters(), ref.parameters()):
        torch.testing.assert_close(safe_get_full_fp32_param(p).cpu(), q.detach(), atol=1e-5, rtol=1e-4)


def test_twin_flow_partial_offload_matches_adamw():
    run_distributed(_twin_flow_worker, 2)


def _grad_introspection_worker(stage):
    import deepspeed_b200 as ds
    from deepspeed_b200.utils import safe_get_full_grad
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(8, 8), torch.nn.Linear(8, 8))
    cfg = {"train_micro_batch_size_per_gpu": 2, "optimizer": {"type": "Adam", "params": {"lr": 1e-2}},
           "zero_optimization": {"stage": stage}}
    eng, *_ = ds.initialize(model=model, config=cfg)
    assert eng.optimizer.fused_in_backward  # gas 1, no clipping: the step is fused into backward
    x = torch.randn(2, 8)
    seen = []
    for _ in range(3):
        eng.backward(eng(x).sum())
        seen.append([safe_get_full_grad(p) for p in model.parameters()])
        eng.step()
    # the first call finds the gradients already consumed and switches to the two-phase step; afterwards they are there
    assert all(g is None for g in seen[0])
    for grads in seen[1:]:
        for g, p in zip(grads, model.parameters()):
            assert g is not None and tuple(g.shape) == tuple(getattr(p, "ds_shape", p.shape)) and g.abs().sum() > 0
    assert not eng.optimizer.fused_in_backward


@pytest.mark.parametrize("stage", [1, 3])
def test_safe_get_full_grad_switches_off_the_fused_step(stage):
    run_distributed(_grad_introspection_worker, 2, (stage, ))
