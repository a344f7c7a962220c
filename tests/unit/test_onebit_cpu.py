"""1-bit collective + optimizers on 2 gloo ranks."""
import pytest
import torch

from tests.common import run_distributed
from tests.unit.simple_model import SimpleModel, base_config, make_batch


def _compressed_allreduce_worker():
    import torch.distributed as td
    from deepspeed_b200.runtime.comm import CompressedBackend
    from deepspeed_b200.runtime.comm.compressed import pack_signs, unpack_signs
    x = torch.randn(64)
    assert torch.equal(unpack_signs(pack_signs(x)), torch.where(x >= 0, 1.0, -1.0))
    be = CompressedBackend()
    r, w = td.get_rank(), td.get_world_size()
    torch.manual_seed(r)
    n = 1000
    padded = n + (8 * w - n % (8 * w)) % (8 * w)
    we, se = torch.zeros(padded), torch.zeros(padded // w)
    # error feedback makes the running mean of the compressed results converge to the true mean
    true_sum, got_sum = torch.zeros(n), torch.zeros(n)
    base = torch.randn(n, generator=torch.Generator().manual_seed(7))
    for it in range(60):
        mine = base + 0.1 * torch.randn(n)
        full = mine.clone()
        td.all_reduce(full)
        true_sum += full / w
        got_sum += be.compressed_allreduce(mine, we, se)
    rel = (got_sum - true_sum).norm() / true_sum.norm()
    assert rel < 0.15, rel


def test_compressed_allreduce():
    run_distributed(_compressed_allreduce_worker, 2)


def _train(opt_name, params, steps):
    import deepspeed_b200 as ds
    torch.manual_seed(0)
    cfg = base_config(0, "fp32", 1, 0.0)
    cfg["optimizer"] = {"type": opt_name, "params": params}
    eng, *_ = ds.initialize(model=SimpleModel(), config=cfg)
    r, w = ds.comm.get_rank(), ds.comm.get_world_size()
    g = torch.Generator().manual_seed(1)
    losses = []
    x, y = make_batch(w, 4, g)  # fixed batch: the loss must go down monotonically-ish
    for it in range(steps):
        loss = eng(x[r * 4:(r + 1) * 4], y[r * 4:(r + 1) * 4])
        eng.backward(loss)
        eng.step()
        losses.append(float(loss.detach()))
    opt = eng.basic_optimizer
    assert opt.freeze_key, "compression stage was never entered"
    assert eng.enable_backward_allreduce is False
    assert losses[-1] < losses[0], losses
    # parameters stay identical across ranks (all communication is symmetric)
    import torch.distributed as td
    from deepspeed_b200.utils import safe_get_full_fp32_param
    for p in eng.module.parameters():
        t = safe_get_full_fp32_param(p).clone()
        ref = t.clone()
        td.broadcast(ref, 0)
        torch.testing.assert_close(t, ref, atol=1e-6, rtol=1e-5)


@pytest.mark.parametrize("name,params", [
    ("OneBitAdam", {"lr": 1e-3, "freeze_step": 10, "eps": 1e-3}),
    ("ZeroOneAdam", {"lr": 1e-3, "eps": 1e-3, "var_freeze_step": 12, "var_update_scaler": 2, "local_step_scaler": 4, "local_step_clipper": 2}),
    ("OneBitLamb", {"lr": 1e-3, "eps": 1e-3, "freeze_step": 10, "max_coeff": 1.0, "min_coeff": 0.01}),
])
def test_onebit_optimizers_train(name, params):
    run_distributed(_train, 2, (name, params, 30))
