"""ZeRO++ (qwZ / hpZ / qgZ) and MiCS on 4 gloo ranks against a plain torch AdamW run."""
import copy

import pytest
import torch

from tests.common import run_distributed
from tests.unit.simple_model import SimpleModel, base_config, make_batch


def _run(zero_extra, tol, steps=4):
    import deepspeed_b200 as ds
    from deepspeed_b200.utils import safe_get_full_fp32_param
    torch.manual_seed(0)
    model = SimpleModel()
    ref = copy.deepcopy(model)
    cfg = base_config(3, "fp32", 1, 0.0)
    cfg["zero_optimization"].update(zero_extra)
    cfg["zero_optimization"]["stage3_param_persistence_threshold"] = 0
    eng, *_ = ds.initialize(model=model, config=cfg)
    r, w = ds.comm.get_rank(), ds.comm.get_world_size()
    ropt = torch.optim.AdamW(ref.parameters(), lr=1e-2, weight_decay=0.01)
    g = torch.Generator().manual_seed(1)
    for it in range(steps):
        x, y = make_batch(w, 4, g)
        loss = eng(x[r * 4:(r + 1) * 4], y[r * 4:(r + 1) * 4])
        eng.backward(loss)
        eng.step()
        ref(x, y).backward()
        ropt.step()
        ropt.zero_grad()
    worst = 0.0
    for (n, p), (_, q) in zip(model.named_parameters(), ref.named_parameters()):
        worst = max(worst, (safe_get_full_fp32_param(p).cpu() - q).abs().max().item())
    assert worst < tol, f"{zero_extra}: max param diff {worst}"
    return eng


def _mics():
    eng = _run({"mics_shard_size": 2}, 1e-5)
    assert eng.optimizer.shard_world == 2 and eng.optimizer.replica_world == 2


def test_mics_shard2_of_4():
    run_distributed(_mics, 4)


def _hpz():
    eng = _run({"zero_hpz_partition_size": 2}, 1e-5)
    assert eng.optimizer.hpz == 2
    assert any(getattr(rt, "sec", None) is not None for rt in eng.optimizer.rts)


def test_hpz_secondary_partition():
    run_distributed(_hpz, 4)


def _quantized():
    # int8 weights / int4 gradients are lossy: parameters track the exact run within quantisation noise
    _run({"zero_quantized_weights": True}, 5e-2)
    _run({"zero_quantized_gradients": True}, 9e-2)  # Adam turns int4 sign flips of tiny grads into O(lr) moves


def test_qwz_qgz():
    run_distributed(_quantized, 2)
