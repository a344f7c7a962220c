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
    _run({"zero_quantized_weights": True}, 6e-2)
    _run({"zero_quantized_gradients": True}, 9e-2)  # Adam turns int4 sign flips of tiny grads into O(lr) moves


def test_qwz_qgz():
    run_distributed(_quantized, 2)


def _mics_hier():
    import torch
    import torch.distributed as td
    from deepspeed_b200.runtime.zero.mics_utils import (_generate_mics_config, create_mics_comm_groups,
                                                        hierarchical_all_gather)
    r = td.get_rank()
    cfg = _generate_mics_config(world_size=8, ndev_per_node=2, shard_size=4)
    assert cfg["shard_groups"] == [[0, 1, 2, 3], [4, 5, 6, 7]] and cfg["replicate_groups"][1] == [1, 5] and cfg["span_nodes"] == 2
    # 4 ranks = one shard group over two 2-GPU "nodes"
    g = create_mics_comm_groups(4, hierarchical_allgather=True, ndev_per_node=2)
    assert g.param_intra_node_group is not None and g.param_inter_node_shard_group is not None
    assert td.get_world_size(g.param_intra_node_group) == 2 and td.get_world_size(g.param_inter_node_shard_group) == 2
    shard = torch.full((3, ), float(r))
    out = torch.empty(12)
    hierarchical_all_gather(out, shard, g)
    assert out.tolist() == [float(i) for i in range(4) for _ in range(3)]


def test_mics_hierarchical_all_gather():
    run_distributed(_mics_hier, 4)


def _qgz_uneven():
    """qgZ with tensor sizes that are not multiples of 8 / of the world size: per-rank chunks are padded to whole quantisation
    groups (the device kernels need groups of a multiple of 8 elements) and the result is still each rank's slice of the mean."""
    import math
    import torch.distributed as td
    from deepspeed_b200.ops.quantizer import quantizer as Q
    from deepspeed_b200.runtime.comm.coalesced_collectives import all_to_all_quant_reduce
    r, w = td.get_rank(), td.get_world_size()
    assert Q.aligned_group_size(5000) == 2048 and Q.aligned_group_size(13) == 16 and Q.aligned_group_size(3) == 8
    for n in (1003, 37, 4096 + 5, 3):
        torch.manual_seed(100 + r)
        x = torch.randn(n)
        (mine, ) = all_to_all_quant_reduce([x.clone()], {"local": None}, num_bits=8)
        full = x.clone()
        td.all_reduce(full)
        full /= w
        per = math.ceil(n / w)
        want = full[r * per:min(n, (r + 1) * per)]
        assert mine.shape == want.shape, (n, mine.shape, want.shape)
        if want.numel():
            assert (mine - want).abs().max() < 0.05 * (full.abs().max() + 1e-6), n


def test_qgz_uneven_sizes_pad_to_aligned_groups():
    run_distributed(_qgz_uneven, 2)


def _loco_worker():
    """LoCo error feedback: the mean of the quantised results over many steps converges to the true mean (the error is
    re-injected), while plain 4-bit qgZ keeps a bias; the error buffer follows the padded chunk layout."""
    import torch.distributed as td
    from deepspeed_b200.ops.quantizer import quantizer as Q
    from deepspeed_b200.runtime.comm.coalesced_collectives import all_to_all_loco_quant_reduce, all_to_all_quant_reduce
    r, w = td.get_rank(), td.get_world_size()
    # single-tensor semantics of the fused op vs its definition
    x = torch.randn(64)
    err = torch.randn(64) * 0.01
    e0 = err.clone()
    q, p = Q.loco_quantize(x, err, 4, num_bits=8, beta=0.5)
    comp = x + e0
    deq = Q.dequantize(q, p, 4, 8, Q.Symmetric, dtype=torch.float32)
    torch.testing.assert_close(err, 0.5 * e0 + 0.5 * (comp - deq))
    n = 1003
    base = torch.randn(n, generator=torch.Generator().manual_seed(7 + r))
    holder = torch.nn.Parameter(base.clone())
    full = base.clone()
    td.all_reduce(full)
    full /= w
    per = -(-n // w)
    want = full[r * per:min(n, (r + 1) * per)]
    acc_loco, acc_plain = torch.zeros_like(want), torch.zeros_like(want)
    T = 40
    for _ in range(T):
        holder.grad = base.clone()
        acc_loco += all_to_all_loco_quant_reduce([holder], {"local": None}, {"err_beta": 0.0, "reset_T": 10**6}, num_bits=4)[0]
        acc_plain += all_to_all_quant_reduce([base.clone()], {"local": None}, num_bits=4)[0]
    e_loco = (acc_loco / T - want).abs().mean().item()
    e_plain = (acc_plain / T - want).abs().mean().item()
    assert e_loco < 0.35 * e_plain + 1e-6, (e_loco, e_plain)
    assert holder.intra_ef_buf[0].numel() % 8 == 0 and holder.intra_ef_buf[1] == T


def test_loco_error_feedback_converges():
    run_distributed(_loco_worker, 2)
