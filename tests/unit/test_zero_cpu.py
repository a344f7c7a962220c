"""ZeRO stages 0-3 on the host tier (gloo, world_size 1 and 2) against a plain torch AdamW run.

Model for the strategy: reference tests/unit/runtime/zero/test_zero.py (tiny models, loss / param
equality across stages) executed by N processes on one host.
"""
import copy

import pytest
import torch

from tests.common import run_distributed
from tests.unit.simple_model import SimpleModel, base_config, make_batch


def _train_and_compare(stage, dtype, gas, clip, steps, tol):
    import deepspeed_b200 as ds
    from deepspeed_b200.utils import safe_get_full_fp32_param
    torch.manual_seed(0)
    model = SimpleModel()
    ref = copy.deepcopy(model)
    eng, _, _, _ = ds.initialize(model=model, config=base_config(stage, dtype, gas, clip))
    r, w = ds.comm.get_rank(), ds.comm.get_world_size()
    ropt = torch.optim.AdamW(ref.parameters(), lr=1e-2, weight_decay=0.01)
    g = torch.Generator().manual_seed(1)
    for it in range(steps * gas):
        x, y = make_batch(w, 4, g)
        xl, yl = x[r * 4:(r + 1) * 4], y[r * 4:(r + 1) * 4]
        if dtype == "bf16":
            xl = xl.bfloat16()
        loss = eng(xl, yl)
        eng.backward(loss)
        eng.step()
        (ref(x, y) / gas).backward()
        if (it + 1) % gas == 0:
            if clip > 0:
                torch.nn.utils.clip_grad_norm_(ref.parameters(), clip)
            ropt.step()
            ropt.zero_grad()
    assert eng.global_steps == steps
    worst = 0.0
    for (n, p), (_, q) in zip(model.named_parameters(), ref.named_parameters()):
        worst = max(worst, (safe_get_full_fp32_param(p).cpu() - q).abs().max().item())
    assert worst < tol, f"stage {stage} {dtype} gas {gas} clip {clip}: max param diff {worst}"


@pytest.mark.parametrize("stage", [0, 1, 2, 3])
def test_single_process_matches_torch(stage):
    run_distributed(_train_and_compare, 1, (stage, "fp32", 1, 0.0, 4, 1e-5))


@pytest.mark.parametrize("stage", [0, 1, 2, 3])
def test_two_ranks_match_torch(stage):
    run_distributed(_train_and_compare, 2, (stage, "fp32", 1, 0.0, 4, 1e-5))


@pytest.mark.parametrize("stage", [1, 3])
def test_grad_accumulation_and_clipping(stage):
    run_distributed(_train_and_compare, 2, (stage, "fp32", 2, 0.5, 3, 1e-5))


def test_bf16_stage3_tracks_fp32_reference():
    # bf16 forward noise + Adam's normalised update => loose tolerance, but must stay in the same basin
    run_distributed(_train_and_compare, 2, (3, "bf16", 1, 0.0, 4, 0.1))


def _fp16_overflow_skips():
    import deepspeed_b200 as ds
    torch.manual_seed(0)
    model = SimpleModel()
    cfg = base_config(2, "fp16")
    cfg["fp16"]["initial_scale_power"] = 4
    eng, _, _, _ = ds.initialize(model=model, config=cfg)
    before = [p.detach().clone() for p in model.parameters()]
    x = torch.full((4, 8), float("inf"), dtype=torch.float16)
    y = torch.zeros(4, dtype=torch.long)
    loss = eng(x, y)
    eng.backward(loss)
    eng.step()
    assert eng.skipped_steps == 1 and eng.optimizer.overflow
    assert eng.optimizer.cur_scale == 2**4 / 2 or eng.optimizer.loss_scaler.cur_hysteresis < 2
    for b, p in zip(before, model.parameters()):
        assert torch.equal(b, p.detach()), "parameters changed on an overflow step"


def test_fp16_overflow_skips_step():
    run_distributed(_fp16_overflow_skips, 2)


def _client_optimizer_and_scheduler():
    import deepspeed_b200 as ds
    torch.manual_seed(0)
    model = SimpleModel()
    ref = copy.deepcopy(model)
    opt = torch.optim.RMSprop(model.parameters(), lr=1e-3)  # not a fused type -> adapter path
    sched = torch.optim.lr_scheduler.StepLR(opt, step_size=2, gamma=0.5)
    cfg = {"train_micro_batch_size_per_gpu": 4, "zero_optimization": {"stage": 2}, "zero_allow_untested_optimizer": True}
    eng, zopt, _, s = ds.initialize(model=model, optimizer=opt, lr_scheduler=sched, config=cfg)
    ropt = torch.optim.RMSprop(ref.parameters(), lr=1e-3)
    rs = torch.optim.lr_scheduler.StepLR(ropt, step_size=2, gamma=0.5)
    r, w = ds.comm.get_rank(), ds.comm.get_world_size()
    g = torch.Generator().manual_seed(3)
    from deepspeed_b200.utils import safe_get_full_fp32_param
    for _ in range(4):
        x, y = make_batch(w, 4, g)
        loss = eng(x[r * 4:(r + 1) * 4], y[r * 4:(r + 1) * 4])
        eng.backward(loss)
        eng.step()
        ref(x, y).backward()
        ropt.step()
        ropt.zero_grad()
        rs.step()
    assert abs(eng.get_lr()[0] - ropt.param_groups[0]["lr"]) < 1e-12
    for p, q in zip(model.parameters(), ref.parameters()):
        assert (safe_get_full_fp32_param(p).cpu() - q).abs().max() < 1e-5


def test_client_optimizer_adapter_and_scheduler():
    run_distributed(_client_optimizer_and_scheduler, 2)


def _gathered_parameters_roundtrip():
    import deepspeed_b200 as ds
    from deepspeed_b200.runtime.zero import GatheredParameters
    torch.manual_seed(0)
    model = SimpleModel()
    eng, _, _, _ = ds.initialize(model=model, config=base_config(3))
    w = model.layers[1].a.weight
    if ds.comm.get_world_size() > 1:
        assert w.numel() == 0, "ZeRO-3 parameter should be released outside forward"
    with GatheredParameters([w], modifier_rank=0):
        assert w.shape == (32, 32)
        if ds.comm.get_rank() == 0:
            w.data.fill_(0.25)
    from deepspeed_b200.utils import safe_get_full_fp32_param
    assert torch.allclose(safe_get_full_fp32_param(w).cpu(), torch.full((32, 32), 0.25))
    x, y = make_batch(1, 4, torch.Generator().manual_seed(0))
    loss = eng(x, y)
    eng.backward(loss)
    eng.step()
    assert torch.isfinite(loss)


def test_gathered_parameters_modify_and_write_back():
    run_distributed(_gathered_parameters_roundtrip, 2)


def _zero_init_context():
    import deepspeed_b200 as ds
    from deepspeed_b200.runtime import zero
    torch.manual_seed(0)
    with zero.Init():
        model = SimpleModel()
    for p in model.parameters():
        assert hasattr(p, "ds_tensor") and p.numel() == 0 and p.ds_numel > 0
    eng, _, _, _ = ds.initialize(model=model, config=base_config(3))
    r, w = ds.comm.get_rank(), ds.comm.get_world_size()
    g = torch.Generator().manual_seed(5)
    losses = []
    for _ in range(5):
        x, y = make_batch(w, 4, g)
        loss = eng(x[r * 4:(r + 1) * 4], y[r * 4:(r + 1) * 4])
        eng.backward(loss)
        eng.step()
        losses.append(loss.item())
    assert all(l == l for l in losses)


def test_zero_init_context_shards_at_construction():
    run_distributed(_zero_init_context, 2)


def _unused_parameter_and_frozen():
    import deepspeed_b200 as ds

    class M(SimpleModel):

        def __init__(self):
            super().__init__()
            self.unused = torch.nn.Linear(4, 4)

    torch.manual_seed(0)
    model = M()
    for p in model.layers[0].parameters():
        p.requires_grad_(False)
    eng, _, _, _ = ds.initialize(model=model, config=base_config(3))
    frozen_before = None
    from deepspeed_b200.utils import safe_get_full_fp32_param
    frozen_before = safe_get_full_fp32_param(model.layers[0].a.weight).clone()
    r, w = ds.comm.get_rank(), ds.comm.get_world_size()
    g = torch.Generator().manual_seed(5)
    for _ in range(3):
        x, y = make_batch(w, 4, g)
        loss = eng(x[r * 4:(r + 1) * 4], y[r * 4:(r + 1) * 4])
        eng.backward(loss)
        eng.step()
    assert torch.equal(frozen_before, safe_get_full_fp32_param(model.layers[0].a.weight))


def test_unused_and_frozen_parameters():
    run_distributed(_unused_parameter_and_frozen, 2)


def _llama_transient_units(stage, ckpt):
    """Tiny Llama with every unit transient (threshold 0): exercises the LM-head / embedding fetch paths,
    the direct flat-gradient writes and recompute-in-backward under ZeRO-3."""
    import deepspeed_b200 as ds
    from deepspeed_b200.models.llama import LlamaForCausalLM, llama_config
    from deepspeed_b200.utils import safe_get_full_fp32_param
    torch.manual_seed(0)
    cfg = llama_config("tiny", checkpoint_layers=ckpt)
    m = LlamaForCausalLM(cfg)
    ref = LlamaForCausalLM(cfg)
    ref.load_state_dict(m.state_dict())
    conf = {"train_micro_batch_size_per_gpu": 2, "optimizer": {"type": "AdamW", "params": {"lr": 1e-3, "weight_decay": 0.1}},
            "zero_optimization": {"stage": stage, "stage3_param_persistence_threshold": 0}}
    eng, _, _, _ = ds.initialize(model=m, config=conf)
    r, w = ds.comm.get_rank(), ds.comm.get_world_size()
    ropt = torch.optim.AdamW(ref.parameters(), lr=1e-3, weight_decay=0.1)
    g = torch.Generator().manual_seed(1)
    for _ in range(3):
        ids = torch.randint(0, cfg.vocab_size, (2 * w, 32), generator=g)
        loss = eng(ids[r * 2:(r + 1) * 2], labels=ids[r * 2:(r + 1) * 2])
        eng.backward(loss)
        eng.step()
        rl = sum(ref(ids[k * 2:(k + 1) * 2], labels=ids[k * 2:(k + 1) * 2]) for k in range(w)) / w
        rl.backward()
        ropt.step()
        ropt.zero_grad()
    worst = max((safe_get_full_fp32_param(p).cpu() - q).abs().max().item()
                for p, q in zip(m.parameters(), ref.parameters()))
    assert worst < 5e-5, worst


@pytest.mark.parametrize("ckpt", [0, 2])
def test_llama_zero3_all_units_transient(ckpt):
    run_distributed(_llama_transient_units, 2, (3, ckpt))
