"""Pipeline parallelism on the host tier: schedule / topology unit tests + 2-stage and (pp=2, dp=2) training
parity against a single-process run (strategy: reference tests/unit/runtime/pipe/*)."""
import copy

import pytest
import torch
from torch import nn

from tests.common import run_distributed


def test_topology_rank_math():
    from deepspeed_b200.runtime.pipe.topology import PipeDataParallelTopology, PipeModelDataParallelTopology, ProcessTopology
    t = ProcessTopology(axes=["a", "b"], dims=[2, 3])
    assert t.world_size() == 6 and t.get_rank(a=1, b=2) == 5 and t.get_coord(4).a == 1 and t.get_coord(4).b == 1
    assert t.get_axis_comm_lists("a") == [[0, 3], [1, 4], [2, 5]]
    assert t.get_axis_comm_lists("b") == [[0, 1, 2], [3, 4, 5]]
    assert t.filter_match(a=0) == [0, 1, 2] and t.get_axis_list("b", 1) == [1, 4]
    t3 = PipeModelDataParallelTopology(num_pp=2, num_mp=2, num_dp=2)
    assert t3.get_rank(pipe=1, data=0, model=1) == 5
    assert t3.get_rank_repr(5) == "model_01"
    assert PipeDataParallelTopology(2, 2).get_axis_comm_lists("data") == [[0, 1], [2, 3]]


@pytest.mark.parametrize("stages,micro", [(1, 1), (2, 4), (4, 4), (4, 8), (3, 2)])
def test_train_schedule_is_consistent(stages, micro):
    from deepspeed_b200.runtime.pipe import schedule as S
    streams = [[c for step in S.TrainSchedule(micro, stages, s) for c in step] for s in range(stages)]
    for s, cmds in enumerate(streams):
        fw = [c.buffer_id for c in cmds if isinstance(c, S.ForwardPass)]
        bw = [c.buffer_id for c in cmds if isinstance(c, S.BackwardPass)]
        assert len(fw) == micro and len(bw) == micro
        assert isinstance(cmds[-1], S.OptimizerStep)
        # live activations never exceed the buffer count
        live, peak = 0, 0
        for c in cmds:
            if isinstance(c, S.ForwardPass):
                live += 1
            elif isinstance(c, S.BackwardPass):
                live -= 1
            peak = max(peak, live)
        assert peak <= S.TrainSchedule(micro, stages, s).num_pipe_buffers()
        n_send = sum(isinstance(c, S.SendActivation) for c in cmds)
        n_recv = sum(isinstance(c, S.RecvGrad) for c in cmds)
        assert n_send == (micro if s < stages - 1 else 0) and n_recv == n_send
    # simulate blocking rendezvous on every link with the executor's exchange fusion: must run to completion
    ops = []
    for s, cmds in enumerate(streams):
        seq = []
        for c in S.fuse_exchanges(cmds):
            if isinstance(c, S.SendActivation):
                seq.append(("send", s + 1, "act"))
            elif isinstance(c, S.RecvActivation):
                seq.append(("recv", s - 1, "act"))
            elif isinstance(c, S.SendGrad):
                seq.append(("send", s - 1, "grad"))
            elif isinstance(c, S.RecvGrad):
                seq.append(("recv", s + 1, "grad"))
            elif isinstance(c, S.SendActivationRecvGrad):
                seq.append(("xchg", s + 1, "act", "grad"))
            elif isinstance(c, S.SendGradRecvActivation):
                seq.append(("xchg", s - 1, "grad", "act"))
        ops.append(seq)
    # every op is a set of sub-ops (an exchange posts its send and its recv together); a stage advances
    # when all sub-ops of its current op have been matched by the neighbour's *current* op
    def subops(op):
        if op[0] == "xchg":
            return [("send", op[1], op[2]), ("recv", op[1], op[3])]
        return [op]

    ptr = [0] * stages
    pending = [subops(ops[s][0]) if ops[s] else [] for s in range(stages)]
    progressed = True
    while progressed:
        progressed = False
        for s in range(stages - 1):
            for a in list(pending[s]):
                for b in list(pending[s + 1]):
                    if a[1] == s + 1 and b[1] == s and a[2] == b[2] and {a[0], b[0]} == {"send", "recv"}:
                        pending[s].remove(a)
                        pending[s + 1].remove(b)
                        progressed = True
                        break
        for s in range(stages):
            while ptr[s] < len(ops[s]) and not pending[s]:
                ptr[s] += 1
                pending[s] = subops(ops[s][ptr[s]]) if ptr[s] < len(ops[s]) else []
                progressed = True
    assert all(ptr[s] >= len(ops[s]) for s in range(stages)), f"pipeline would dead-lock: {ptr} vs {[len(o) for o in ops]}"


class _Blk(nn.Module):

    def __init__(self, d):
        super().__init__()
        self.l = nn.Linear(d, d)

    def forward(self, x):
        return torch.tanh(self.l(x))


def _pipe_train(num_stages, micro=4):
    import deepspeed_b200 as ds
    from deepspeed_b200.pipe import LayerSpec, PipelineModule
    torch.manual_seed(0)
    d, L, mbs = 16, 4, 2
    ref_layers = [_Blk(d) for _ in range(L)]
    ref = nn.Sequential(*copy.deepcopy(ref_layers))
    w = torch.distributed.get_world_size()
    dp = w // num_stages
    model = PipelineModule(layers=copy.deepcopy(ref_layers), num_stages=num_stages, loss_fn=nn.MSELoss(),
                           partition_method="uniform")
    cfg = {"train_micro_batch_size_per_gpu": mbs, "gradient_accumulation_steps": micro,
           "optimizer": {"type": "SGD", "params": {"lr": 0.1}}, "zero_optimization": {"stage": 0}}
    eng, _, _, _ = ds.initialize(model=model, config=cfg)
    dp_rank = eng.grid.get_data_parallel_id()
    g = torch.Generator().manual_seed(5)
    ropt = torch.optim.SGD(ref.parameters(), lr=0.1)
    for it in range(3):
        xs = torch.randn(dp, micro, mbs, d, generator=g)
        ys = torch.randn(dp, micro, mbs, d, generator=g)
        data = iter([(xs[dp_rank, m], ys[dp_rank, m]) for m in range(micro)])
        loss = eng.train_batch(data_iter=data)
        total = 0.0
        for r in range(dp):
            for m in range(micro):
                l = nn.functional.mse_loss(ref(xs[r, m]), ys[r, m]) / (micro * dp)
                l.backward()
                total += l.item()
        ropt.step()
        ropt.zero_grad()
        assert abs(loss.item() - total) < 1e-5, (loss.item(), total)
    # compare the layers this stage owns
    for idx in range(model._local_start, model._local_stop):
        mine = dict(model.named_parameters())[f"{idx}.l.weight"]
        from deepspeed_b200.utils import safe_get_full_fp32_param
        assert (safe_get_full_fp32_param(mine).cpu() - ref[idx].l.weight).abs().max() < 1e-5


def test_pipeline_two_stages_matches_sequential():
    run_distributed(_pipe_train, 2, (2, ))


def test_pipeline_single_micro_batch_per_step():
    """One micro batch per step: the activation send and the gradient receive of the SAME micro batch are a dependency,
    not an exchange (they must not be posted as one grouped operation)."""
    from deepspeed_b200.runtime.pipe import schedule as S
    fused = S.fuse_exchanges([S.SendActivation(buffer_id=0), S.RecvGrad(buffer_id=0), S.SendActivation(buffer_id=1),
                              S.RecvGrad(buffer_id=0)])
    assert [type(c).__name__ for c in fused] == ["SendActivation", "RecvGrad", "SendActivationRecvGrad"]
    run_distributed(_pipe_train, 2, (2, 1))


def test_pipeline_pp2_dp2_matches_sequential():
    run_distributed(_pipe_train, 4, (2, ), timeout=300)


def _raw_p2p():
    import torch
    import deepspeed_b200 as ds
    from deepspeed_b200 import comm as dist
    from deepspeed_b200.runtime.pipe import p2p
    from deepspeed_b200.runtime.pipe.topology import PipeDataParallelTopology, PipelineParallelGrid
    ds.init_distributed()
    grid = PipelineParallelGrid(PipeDataParallelTopology(num_pp=2, num_dp=1))
    p2p.init_process_groups(grid)
    assert p2p.can_send_recv()
    if grid.get_stage_id() == 0:
        p2p.send(torch.arange(6.0), 1)
        p2p.send(torch.ones(3) * 7, 1, async_op=True)
        p2p.wait()
    else:
        a, b = torch.zeros(6), torch.zeros(3)
        p2p.recv(a, 0)
        p2p.recv(b, 0, async_op=True)
        p2p.wait()
        assert torch.equal(a, torch.arange(6.0)) and torch.equal(b, torch.ones(3) * 7)


def test_raw_p2p_send_recv():
    run_distributed(_raw_p2p, 2)


class _Emb(nn.Module):

    def __init__(self, v, d):
        super().__init__()
        self.weight = nn.Parameter(torch.randn(v, d) * 0.1)

    def forward(self, ids):
        return torch.nn.functional.embedding(ids, self.weight)


def _tied_head(mod, x):
    return torch.matmul(x, mod.weight.t())


def _pipe_variants(mode):
    """ZeRO-1 under the pipeline engine, per-layer activation checkpointing, and an embedding tied between the first and the
    last stage (``TiedLayerSpec``: the copies must receive the same summed gradient and stay identical)."""
    import deepspeed_b200 as ds
    from deepspeed_b200.pipe import LayerSpec, PipelineModule, TiedLayerSpec
    from deepspeed_b200.utils import safe_get_full_fp32_param
    torch.manual_seed(0)
    d, micro, mbs, V = 16, 2, 2, 32
    if mode == "tied":
        layers = [TiedLayerSpec("emb", _Emb, V, d), LayerSpec(_Blk, d), LayerSpec(_Blk, d),
                  TiedLayerSpec("emb", _Emb, V, d, forward_fn=_tied_head)]
        loss_fn = lambda logits, y: nn.functional.cross_entropy(logits.reshape(-1, V), y.reshape(-1))
    else:
        layers = [_Blk(d) for _ in range(4)]
        loss_fn = nn.MSELoss()
    kw = {"activation_checkpoint_interval": 1} if mode == "ckpt" else {}
    model = PipelineModule(layers=layers, num_stages=2, loss_fn=loss_fn, partition_method="uniform", **kw)
    cfg = {"train_micro_batch_size_per_gpu": mbs, "gradient_accumulation_steps": micro,
           "optimizer": {"type": "Adam", "params": {"lr": 1e-2}}, "zero_optimization": {"stage": 1 if mode == "zero1" else 0}}
    eng, *_ = ds.initialize(model=model, config=cfg)
    losses = []
    for _ in range(6):
        g = torch.Generator().manual_seed(1)
        if mode == "tied":
            ids = torch.randint(0, V, (mbs, 5), generator=g)
            data = [(ids, ids) for _ in range(micro)]
        else:
            data = [(torch.randn(mbs, d, generator=g), torch.randn(mbs, d, generator=g)) for _ in range(micro)]
        losses.append(eng.train_batch(data_iter=iter(data)).item())
    assert losses[-1] < losses[0] - 0.1, losses
    if mode == "tied":
        w = [p for n, p in model.named_parameters() if p.shape == (V, d)][0]
        full = safe_get_full_fp32_param(w)
        both = [torch.empty_like(full) for _ in range(2)]
        torch.distributed.all_gather(both, full)
        assert torch.allclose(both[0], both[1], atol=1e-6), (both[0] - both[1]).abs().max()
    return losses


@pytest.mark.parametrize("mode", ["zero1", "ckpt", "tied"])
def test_pipeline_zero1_checkpointing_and_tied_layers(mode):
    run_distributed(_pipe_variants, 2, (mode, ))
