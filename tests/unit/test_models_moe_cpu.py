"""GPT-2 ZeRO-1 on gloo (BASELINE config 1) and Mixtral-style MoE with expert parallelism on the host tier."""
import copy

import pytest
import torch

from tests.common import run_distributed


def _gpt2_zero1():
    import deepspeed_b200 as ds
    from deepspeed_b200.models.gpt2 import GPT2LMHeadModel, gpt2_config
    torch.manual_seed(0)
    cfg = gpt2_config("gpt2-tiny")
    model = GPT2LMHeadModel(cfg)
    ref = copy.deepcopy(model)
    eng, _, _, _ = ds.initialize(model=model, config={
        "train_micro_batch_size_per_gpu": 2, "optimizer": {"type": "AdamW", "params": {"lr": 1e-3, "weight_decay": 0.01}},
        "zero_optimization": {"stage": 1}, "scheduler": {"type": "WarmupLR", "params": {"warmup_num_steps": 4, "warmup_max_lr": 1e-3}}})
    r, w = ds.comm.get_rank(), ds.comm.get_world_size()
    ropt = torch.optim.AdamW(ref.parameters(), lr=1e-3, weight_decay=0.01)
    from deepspeed_b200.runtime.lr_schedules import WarmupLR
    rs = WarmupLR(ropt, warmup_num_steps=4, warmup_max_lr=1e-3)
    g = torch.Generator().manual_seed(3)
    losses = []
    for _ in range(4):
        ids = torch.randint(0, cfg.vocab_size, (2 * w, 24), generator=g)
        loss = eng(ids[r * 2:(r + 1) * 2], labels=ids[r * 2:(r + 1) * 2])
        eng.backward(loss)
        eng.step()
        losses.append(loss.item())
        rl = sum(ref(ids[k * 2:(k + 1) * 2], labels=ids[k * 2:(k + 1) * 2]) for k in range(w)) / w
        rl.backward()
        ropt.step()
        ropt.zero_grad()
        rs.step()
    from deepspeed_b200.utils import safe_get_full_fp32_param
    worst = max((safe_get_full_fp32_param(p).cpu() - q).abs().max().item()
                for p, q in zip(model.parameters(), ref.parameters()))
    assert worst < 1e-4, worst
    assert model.lm_head.weight.data_ptr() == model.wte.weight.data_ptr(), "tied weights must stay tied"


def test_gpt2_small_arch_zero1_gloo_ws2():
    run_distributed(_gpt2_zero1, 2)


def _moe_ep(stage):
    import deepspeed_b200 as ds
    from deepspeed_b200.models.mixtral import MixtralForCausalLM, mixtral_config
    torch.manual_seed(0)
    w = torch.distributed.get_world_size()
    cfg = mixtral_config("tiny-moe", ep_size=w)
    model = MixtralForCausalLM(cfg)
    eng, opt, _, _ = ds.initialize(model=model, config={
        "train_micro_batch_size_per_gpu": 2, "optimizer": {"type": "AdamW", "params": {"lr": 2e-3}},
        "zero_optimization": {"stage": stage}, "gradient_clipping": 1.0})
    from deepspeed_b200.runtime.zero.multi import ZeroOptimizerGroup
    assert isinstance(opt, ZeroOptimizerGroup) and len(opt.parts) == 2
    r = ds.comm.get_rank()
    g = torch.Generator().manual_seed(11)
    ids = torch.randint(0, cfg.vocab_size, (2 * w, 32), generator=g)[r * 2:(r + 1) * 2]
    losses = []
    for _ in range(6):
        loss = eng(ids, labels=ids)
        eng.backward(loss)
        eng.step()
        losses.append(loss.item())
    assert all(l == l for l in losses) and losses[-1] < losses[0], losses
    # the dense (router / attention) weights must be identical on both ranks, experts differ
    from deepspeed_b200.utils import safe_get_full_fp32_param
    router = model.layers[0].block_sparse_moe.deepspeed_moe.gate.wg.weight
    full = safe_get_full_fp32_param(router)
    both = [torch.empty_like(full) for _ in range(w)]
    torch.distributed.all_gather(both, full)
    assert torch.equal(both[0], both[-1])


@pytest.mark.parametrize("stage", [1, 2])
def test_mixtral_style_moe_expert_parallel_ws2(stage):
    run_distributed(_moe_ep, 2, (stage, ))


def test_gate_drop_policies_and_dense_view():
    """``topkgating`` honours ``drop_policy`` (most probable assignments survive vs first come first served) and its
    result unpacks into the reference's dense ``(l_aux, combine_weights, dispatch_mask, exp_counts)`` form."""
    import torch
    from deepspeed_b200.moe.sharded_moe import topkgating
    logits = torch.tensor([[0.11, 0.2, 0.1, 0.3], [0.3, 0.4, 0.11, 0.1], [0.11, 0.1, 0.6, 0.5], [0.1, 0.11, 0.7, 0.8]])

    def dense(truth, cap=2):
        t = torch.zeros(4, 4, cap)
        i, j, k = torch.tensor(truth).t()
        t[i, j, k] = 1
        return t

    probs = topkgating(logits, 2, 1, min_capacity=1, drop_policy="probs")
    assert torch.equal(dense([[0, 1, 0], [1, 0, 0], [1, 1, 1], [2, 2, 0], [2, 3, 0], [3, 2, 1], [3, 3, 1]]), probs[2])
    pos = topkgating(logits, 2, 1, min_capacity=1, drop_policy="position")
    assert torch.equal(dense([[0, 1, 0], [0, 3, 0], [1, 0, 0], [1, 1, 1], [2, 2, 0], [2, 3, 1], [3, 2, 1]]), pos[2])
    l_aux, combine, mask, counts = pos
    assert combine.shape == (4, 4, 2) and torch.equal(combine != 0, mask) and counts.tolist() == [1, 2, 2, 3]
    assert l_aux.ndim == 0
