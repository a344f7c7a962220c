"""Hybrid engine: train with ZeRO-3 (2 gloo ranks), generate from the *current* weights, train again."""
import torch

from tests.common import run_distributed


def _worker():
    import deepspeed_b200 as ds
    from deepspeed_b200.models.llama import LlamaConfig, LlamaForCausalLM
    torch.manual_seed(0)
    cfg = LlamaConfig(vocab_size=64, hidden_size=32, intermediate_size=64, num_hidden_layers=2, num_attention_heads=4,
                      num_key_value_heads=2, max_position_embeddings=64)
    model = LlamaForCausalLM(cfg).float()
    eng, *_ = ds.initialize(model=model, config={
        "train_micro_batch_size_per_gpu": 2, "optimizer": {"type": "AdamW", "params": {"lr": 1e-2}},
        "zero_optimization": {"stage": 3, "stage3_param_persistence_threshold": 0},
        "hybrid_engine": {"enabled": True, "max_out_tokens": 32, "release_inference_cache": False}})
    from deepspeed_b200.runtime.hybrid_engine import DeepSpeedHybridEngine
    assert isinstance(eng, DeepSpeedHybridEngine)
    r = ds.comm.get_rank()
    g = torch.Generator().manual_seed(r)
    prompt = torch.randint(0, 64, (2, 5), generator=torch.Generator().manual_seed(9))

    def ref_generate():
        from deepspeed_b200.runtime.zero.partition_parameters import GatheredParameters
        with GatheredParameters(list(model.parameters())):
            return model.generate_greedy(prompt, max_new_tokens=4)

    out0 = eng.generate(prompt, max_new_tokens=4)
    assert torch.equal(out0, ref_generate())
    for _ in range(3):
        ids = torch.randint(0, 64, (2, 12), generator=g)
        loss = eng(ids, labels=ids)
        loss = loss[0] if isinstance(loss, tuple) else loss
        eng.backward(loss)
        eng.step()
    out1 = eng.generate(prompt, max_new_tokens=4)     # re-packed from the updated shards
    assert torch.equal(out1, ref_generate())
    assert eng._packed_at_step == 3
    rep = eng.get_latency_report()
    assert rep["generate_s"] > 0


def test_hybrid_engine_zero3_ws2():
    run_distributed(_worker, 2, timeout=300)


def _tp_worker():
    """Generation-time tensor parallelism: the two ranks shard the packed inference weights / KV cache, serve the
    union of their prompts as one batch, and each gets its own continuation back (reference hybrid_engine.py:168-205)."""
    import deepspeed_b200 as ds
    from deepspeed_b200.models.llama import LlamaConfig, LlamaForCausalLM
    torch.manual_seed(0)
    cfg = LlamaConfig(vocab_size=64, hidden_size=32, intermediate_size=64, num_hidden_layers=2, num_attention_heads=4,
                      num_key_value_heads=2, max_position_embeddings=64)
    model = LlamaForCausalLM(cfg).float()
    eng, *_ = ds.initialize(model=model, config={
        "train_micro_batch_size_per_gpu": 2, "optimizer": {"type": "AdamW", "params": {"lr": 1e-2}},
        "zero_optimization": {"stage": 3, "stage3_param_persistence_threshold": 0},
        "hybrid_engine": {"enabled": True, "max_out_tokens": 32, "inference_tp_size": 2}})
    r = ds.comm.get_rank()
    prompt = torch.randint(0, 64, (2, 5), generator=torch.Generator().manual_seed(100 + r))   # rank-specific prompts

    def ref_generate():
        from deepspeed_b200.runtime.zero.partition_parameters import GatheredParameters
        with GatheredParameters(list(model.parameters())):
            return model.generate_greedy(prompt, max_new_tokens=4)

    for it in range(2):
        out = eng.generate(prompt, max_new_tokens=4)
        assert out.shape == (2, 9) and torch.equal(out, ref_generate()), (r, it)
        assert eng._ragged._model.tp_size == 2 and eng._ragged._model.hq == 2
        ids = torch.randint(0, 64, (2, 12), generator=torch.Generator().manual_seed(r + it))
        loss = eng(ids, labels=ids)
        loss = loss[0] if isinstance(loss, tuple) else loss
        eng.backward(loss)
        eng.step()


def test_hybrid_engine_inference_tp2_ws2():
    run_distributed(_tp_worker, 2, timeout=300)
