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


def _async_param_gather():
    import torch
    import deepspeed_b200 as ds
    from deepspeed_b200.runtime.zero import partition_parameters as PP
    ds.init_distributed()
    torch.manual_seed(0)
    with ds.zero.Init():
        m = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.Linear(32, 8))
    ref = [PP.materialize_full(p).clone() for p in m.parameters()]
    ps = list(m.parameters())
    assert all(p.data.numel() == 0 for p in ps)
    h = ps[0].all_gather(param_list=ps, async_op=True)
    assert all(p.ds_status == PP.ZeroParamStatus.INFLIGHT for p in ps)
    h.wait()
    h.wait()  # idempotent
    for p, r in zip(ps, ref):
        assert p.ds_status == PP.ZeroParamStatus.AVAILABLE and torch.equal(p.data, r)
        PP.free_param(p)
        assert p.data.numel() == 0 and p.ds_status == PP.ZeroParamStatus.NOT_AVAILABLE
    # int8 quantised gather (ZeRO++ qwZ): close to the exact tensors
    big = [p for p in ps if p.ds_tensor.numel() % 8 == 0]
    hq = PP.all_gather_coalesced(big, quantize=True)
    hq.wait()
    for p in big:
        r = ref[[id(x) for x in ps].index(id(p))]
        assert (p.data.float() - r.float()).abs().max() <= r.abs().max() / 127 + 1e-6


def test_async_param_gather_handles():
    from tests.common import run_distributed
    run_distributed(_async_param_gather, 2)


def test_partition_parameter_helpers():
    import torch
    from deepspeed_b200.runtime.zero import partition_parameters as PP

    class A:
        pass

    class B(A):
        pass

    class C(B):
        pass

    assert PP.get_all_subclasses(A) == {A, B, C} and PP.get_all_subclasses(A, include_root=False) == {B, C}
    mk = PP.zero_wrapper_for_fp_tensor_constructor(torch.ones, torch.bfloat16)
    assert mk(3).dtype == torch.bfloat16 and mk(3, dtype=torch.int32).dtype == torch.int32
    nt = PP.get_new_tensor_fn_for_dtype(torch.float16)
    assert nt(torch.Tensor, (2, 3)).dtype == torch.float16
    q = PP.CUDAQuantizer()
    g = q._groups_for(32000)
    assert 32000 % (8 * g) == 0 and 32000 / g <= 16000
    x = torch.randn(32000)
    qv, sc = q.quantize(x)
    assert (q.dequantize(qv, sc, dtype=torch.float32) - x).abs().max() <= x.abs().max() / 127 + 1e-6
    assert PP.InsertPostInitMethodToModuleSubClasses is PP.Init


def _zero_conveniences():
    import torch
    import deepspeed_b200 as ds
    model = torch.nn.Linear(8, 8)
    eng, opt, *_ = ds.initialize(model=model, config={"train_batch_size": 2, "optimizer": {"type": "Adam", "params": {"lr": 1e-3}},
                                                      "zero_optimization": {"stage": 2}})
    assert opt.get_lr() == 1e-3
    opt.set_lr(5e-4)
    assert all(g["lr"] == 5e-4 for g in opt.param_groups) and not opt.dynamic_loss_scale and not opt.has_overflow()
    loss = eng(torch.randn(1, 8)).square().mean()
    eng.backward(loss)
    assert opt.get_grad_norm_direct() >= 0
    eng.step()
    opt.override_loss_scale(4.0)
    assert opt.loss_scale == 4.0 and opt.custom_loss_scaler
    ts = [torch.arange(3.0), torch.arange(4.0) + 10]
    flat = type(opt).defragment(ts)
    assert flat.tolist() == [0, 1, 2, 10, 11, 12, 13] and ts[1].data_ptr() == flat[3:].data_ptr()


def test_zero_optimizer_conveniences():
    from tests.common import run_distributed
    run_distributed(_zero_conveniences, 2)
