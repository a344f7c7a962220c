"""Small compatibility utilities: bwc accessors, numa binding maths, MoE token mappings under TP, torch gates."""
import pytest
import torch

from tests.common import run_distributed


def test_bwc_accessors_handle_all_spellings():
    from deepspeed_b200.utils import bwc

    class New:
        get_tensor_model_parallel_rank = staticmethod(lambda: 3)
        get_tensor_model_parallel_world_size = staticmethod(lambda: 4)
        get_pipeline_model_parallel_world_size = staticmethod(lambda: 2)

    class Old:
        get_model_parallel_rank = staticmethod(lambda: 1)
        get_slice_parallel_world_size = staticmethod(lambda: 8)
        get_pipe_parallel_world_size = staticmethod(lambda: 5)

    assert bwc.bwc_tensor_model_parallel_rank(New) == 3 and bwc.bwc_tensor_model_parallel_world_size(New) == 4
    assert bwc.bwc_pipeline_parallel_world_size(New) == 2
    assert bwc.bwc_tensor_model_parallel_rank(Old) == 1 and bwc.bwc_tensor_model_parallel_world_size(Old) == 8
    assert bwc.bwc_pipeline_parallel_world_size(Old) == 5
    assert bwc.bwc_tensor_model_parallel_world_size(None) == 1


def test_numa_ranges_and_binding(monkeypatch):
    from deepspeed_b200.utils import numa
    assert numa.parse_range("4") == [4] and numa.parse_range("2-5") == [2, 3, 4, 5]
    assert numa.parse_range_list("0-2,5,7-8") == [0, 1, 2, 5, 7, 8]
    with pytest.raises(ValueError):
        numa.parse_range_list("3-1")
    with pytest.raises(ValueError):
        numa.parse_range_list("0-4,3")
    monkeypatch.setattr(numa, "get_numa_cores", lambda: [list(range(0, 8)), list(range(8, 16))])
    monkeypatch.delenv("KMP_AFFINITY", raising=False)
    n, cmd = numa.get_numactl_cmd("", 4, 3)
    assert n == 4 and cmd == ["numactl", "-m", "1", "-C", "12-15"]
    n, cmd = numa.get_numactl_cmd("0-3,8-11", 2, 1)
    assert n == 4 and cmd == ["numactl", "-m", "1", "-C", "8-11"]
    monkeypatch.setenv("KMP_AFFINITY", "x")
    with pytest.raises(ValueError):
        numa.get_numactl_cmd("", 1, 0)


def test_torch_gates_and_types():
    from deepspeed_b200.utils.torch import register_grad_hook, required_torch_version
    from deepspeed_b200.utils.types import GATED_ACTIVATION_TYPES, ActivationFuncType
    from deepspeed_b200.utils.config import get_timers_config
    assert required_torch_version(min_version=1.8) and not required_torch_version(max_version=1.8)
    p = torch.nn.Parameter(torch.ones(3))
    seen = []
    register_grad_hook(p, lambda q: seen.append(q.grad.clone()))
    (p * 2).sum().backward()
    assert seen and torch.equal(seen[0], torch.full((3, ), 2.0))
    assert ActivationFuncType.GATED_SILU in GATED_ACTIVATION_TYPES
    assert get_timers_config({"timers": {"throughput": {"enabled": False}}}).enabled is False
    assert get_timers_config({}).synchronized is True


def _tp_moe():
    import torch.distributed as td
    import deepspeed_b200 as ds
    from deepspeed_b200.moe.layer import MoE
    from deepspeed_b200.moe.mappings import drop_tokens, gather_tokens
    from deepspeed_b200.utils import groups
    r = td.get_rank()
    groups.initialize(tp_size=2)  # both ranks form one TP group, dp = 1
    x = torch.arange(24, dtype=torch.float32).reshape(2, 4, 3).requires_grad_(True)
    d = drop_tokens(x, dim=1)
    assert d.shape == (2, 2, 3) and torch.equal(d, x[:, 2 * r:2 * r + 2])
    g = gather_tokens(d, dim=1)
    assert torch.equal(g, x)
    g.sum().backward()
    assert torch.equal(x.grad, torch.ones_like(x))  # gather bwd = drop, drop bwd = gather -> every element once
    # a MoE block under TP=2 with replicated experts must match the single-process result
    torch.manual_seed(0)
    expert = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.GELU(), torch.nn.Linear(16, 8))
    moe = MoE(8, expert, num_experts=2, ep_size=1, k=1, capacity_factor=2.0, min_capacity=4, use_rts=False)
    moe.set_deepspeed_parallelism()
    torch.manual_seed(1)
    inp = torch.randn(6, 8)
    y, _, _ = moe(inp)
    groups_tp = groups._get_model_parallel_world_size()
    assert groups_tp == 2
    # reference result: same layer evaluated with the mappings disabled (expert_tp=True keeps all tokens on every rank)
    moe.deepspeed_moe.expert_tp = True
    y_ref, _, _ = moe(inp)
    assert torch.allclose(y, y_ref, atol=1e-6)
    y.sum().backward()
    assert all(p.grad is not None for p in moe.parameters())


def test_moe_token_mappings_under_tp():
    run_distributed(_tp_moe, 2)


def _cdb():
    import deepspeed_b200.comm as dist
    from deepspeed_b200.comm import comm as C
    from deepspeed_b200.comm.torch import TorchBackend, all_reduce_comm_off
    from deepspeed_b200.comm.utils import get_msg_size_from_args, get_world_size_from_launcher
    dist.init_distributed("gloo")
    cdb = C.cdb
    assert isinstance(cdb, TorchBackend) and cdb.get_world_size() == 2
    t = torch.ones(4) * (cdb.get_rank() + 1)
    cdb.all_reduce(t, op=dist.ReduceOp.SUM)
    assert torch.equal(t, torch.full((4, ), 3.0))
    out = torch.empty(8)
    cdb.all_gather_into_tensor(out, torch.full((4, ), float(cdb.get_rank())))
    assert out.tolist() == [0.0] * 4 + [1.0] * 4
    all_reduce_comm_off(True)
    u = torch.ones(2)
    cdb.all_reduce(u).wait()
    assert torch.equal(u, torch.ones(2))  # switched off: untouched
    all_reduce_comm_off(False)
    assert get_msg_size_from_args(cdb.all_reduce, torch.zeros(10)) == 40
    assert get_msg_size_from_args(cdb.reduce_scatter, torch.zeros(2), [torch.zeros(3), torch.zeros(5)]) == 32
    assert get_world_size_from_launcher() == 2


def test_object_style_backend():
    run_distributed(_cdb, 2)


def test_snip_momentum_block_pruning():
    import torch
    from deepspeed_b200.compression.helper import generate_pruners, register_on_step_begin, rewrite_optimizer_step
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.ReLU(), torch.nn.Linear(32, 8), torch.nn.Linear(8, 3))
    model.pruners = generate_pruners({"target_sparsity": 0.5, "pattern": "4x1", "pruning_frequency": 2, "start_step": 1,
                                      "end_step": 9, "excluded_op_names": [r"^3$"]}, model)
    pr = model.pruners[0]
    assert set(pr.modules) == {"0", "2"}  # the 3x8 head is excluded by name
    h = register_on_step_begin(model)
    opt = rewrite_optimizer_step(torch.optim.SGD(model.parameters(), lr=0.05))
    opt.pruners = model.pruners
    seen = []
    for step in range(14):
        x = torch.randn(8, 16)
        loss = model(x).square().mean()
        opt.zero_grad()
        loss.backward()
        opt.step()
        seen.append(pr.sparsity())
    h.remove()
    assert seen[0] == 0.0 and all(b >= a - 1e-9 for a, b in zip(seen, seen[1:]))  # cubic ramp: monotone
    assert abs(seen[-1] - 0.5) < 0.02
    for name, mod in pr.modules.items():
        w = mod.weight
        assert torch.all(w[~pr.masks[name]] == 0)
        blocks = (w != 0).reshape(w.shape[0] // 4, 4, w.shape[1]).float().mean(1)
        assert torch.all((blocks == 0) | (blocks == 1))  # whole 4x1 blocks live or die together


def test_sparsity_layout_setters_compose_to_make_layout():
    import torch
    from deepspeed_b200.ops.sparse_attention.sparsity_config import (BigBirdSparsityConfig, BSLongformerSparsityConfig,
                                                                     FixedSparsityConfig, VariableSparsityConfig)
    cases = [(FixedSparsityConfig(num_heads=2, block=16, num_local_blocks=4, num_global_blocks=1, attention="unidirectional"),
              ("set_local_layout", "set_global_layout")),
             (BSLongformerSparsityConfig(num_heads=2, block=16, num_sliding_window_blocks=3, global_block_indices=[0, 5]),
              ("set_sliding_window_layout", "set_global_layout")),
             (VariableSparsityConfig(num_heads=2, block=16, local_window_blocks=[2, 4], global_block_indices=[0]),
              ("set_local_layout", "set_global_layout")),
             (BigBirdSparsityConfig(num_heads=2, block=16, num_random_blocks=0, num_sliding_window_blocks=3, num_global_blocks=1),
              ("set_random_layout", "set_sliding_window_layout", "set_global_layout_itc"))]
    for cfg, setters in cases:
        want = cfg.make_layout(256)
        lay = cfg.setup_layout(256)
        for h in range(cfg.num_layout_heads):
            for name in setters:
                lay = getattr(cfg, name)(h, lay)
        assert torch.equal(cfg.check_and_propagate_first_head_layout(lay), want), type(cfg).__name__
