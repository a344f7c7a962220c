"""Ragged engine on host tensors: logits parity with HF for every supported family, scheduling, flush."""
import pytest
import torch

from deepspeed_b200.inference.v2 import build_hf_engine, build_engine_from_model, SchedulingResult
from deepspeed_b200.inference.v2.ragged import BlockedAllocator

transformers = pytest.importorskip("transformers")


def _hf(mt):
    from transformers import AutoConfig, AutoModelForCausalLM
    common = dict(vocab_size=128, hidden_size=64, num_hidden_layers=2, num_attention_heads=4, max_position_embeddings=256)
    kw = {
        "llama": dict(num_key_value_heads=2, intermediate_size=96),
        "mistral": dict(num_key_value_heads=2, intermediate_size=96, sliding_window=None),
        "qwen2": dict(num_key_value_heads=2, intermediate_size=96),
        "phi3": dict(num_key_value_heads=4, intermediate_size=96, pad_token_id=0),
        "mixtral": dict(num_key_value_heads=2, intermediate_size=96, num_local_experts=4, num_experts_per_tok=2),
        "opt": dict(ffn_dim=96, word_embed_proj_dim=64),
        "gpt2": dict(n_embd=64, n_layer=2, n_head=4, n_positions=256),
        "gpt_neox": dict(intermediate_size=96, rotary_pct=0.5),
        "falcon": dict(new_decoder_architecture=False, multi_query=True, parallel_attn=True, bias=False, alibi=False),
        "phi": dict(intermediate_size=96, partial_rotary_factor=0.5),
    }[mt]
    cfg = AutoConfig.for_model(mt, **{**common, **kw})
    torch.manual_seed(0)
    m = AutoModelForCausalLM.from_config(cfg).float().eval()
    return m


def _engine(m, **kw):
    cfg = {"state_manager": {"max_context": 256, "max_ragged_batch_size": 128, "max_ragged_sequence_count": 8,
                             "memory_config": {"mode": "allocate", "size": 32}}}
    e = build_hf_engine(m, cfg, dtype=torch.float32, device="cpu", **kw)
    return e


@pytest.mark.parametrize("mt", ["llama", "mistral", "qwen2", "phi3", "mixtral", "opt", "gpt2", "gpt_neox", "falcon", "phi"])
def test_family_parity(mt):
    m = _hf(mt)
    e = _engine(m)
    torch.manual_seed(1)
    p0 = torch.randint(0, 128, (11, ))
    p1 = torch.randint(0, 128, (5, ))
    logits = e.put([0, 1], [p0, p1])
    with torch.no_grad():
        r0 = m(p0[None]).logits[0, -1]
        r1 = m(p1[None]).logits[0, -1]
    torch.testing.assert_close(logits[0], r0, atol=2e-4, rtol=1e-3)
    torch.testing.assert_close(logits[1], r1, atol=2e-4, rtol=1e-3)
    # one decode step for seq 0 mixed with a continuation chunk for seq 1
    n0 = logits[0].argmax().reshape(1)
    c1 = torch.randint(0, 128, (3, ))
    logits2 = e.put([0, 1], [n0, c1])
    with torch.no_grad():
        r0 = m(torch.cat([p0, n0])[None]).logits[0, -1]
        r1 = m(torch.cat([p1, c1])[None]).logits[0, -1]
    torch.testing.assert_close(logits2[0], r0, atol=2e-4, rtol=1e-3)
    torch.testing.assert_close(logits2[1], r1, atol=2e-4, rtol=1e-3)


def test_scheduling_and_flush():
    e = _engine(_hf("llama"))
    total = int(e.free_blocks[0])
    assert e.can_schedule([0], [300]) == SchedulingResult.SequenceTokenLimitExceeded
    assert e.can_schedule([0], [200]) == SchedulingResult.BatchTokenLimitExceeded
    assert e.can_schedule(list(range(9)), [1] * 9) == SchedulingResult.BatchSequenceLimitExceeded
    e.put([7], [torch.arange(10)])
    assert int(e.free_blocks[0]) == total - 1
    toks, blocks = e.query(7, 128, 4)
    assert toks == 128 and blocks == 1  # 10 + 128 tokens = 2 blocks of 128
    assert e.get_remaining_block_capacity(7) == 118
    e.flush(7)
    assert int(e.free_blocks[0]) == total


def test_blocked_allocator():
    a = BlockedAllocator(8)
    x = a.allocate(3)
    assert a.free_blocks == 5 and len(set(x.tolist())) == 3
    a.free(x[:2])
    assert a.free_blocks == 7
    with pytest.raises(ValueError):
        a.free(x[:1])
    with pytest.raises(ValueError):
        a.allocate(9)


def test_from_b200_llama_and_quantized():
    from deepspeed_b200.models.llama import LlamaConfig, LlamaForCausalLM
    cfg = LlamaConfig(vocab_size=128, hidden_size=64, intermediate_size=96, num_hidden_layers=2, num_attention_heads=4,
                      num_key_value_heads=2, max_position_embeddings=256)
    torch.manual_seed(0)
    m = LlamaForCausalLM(cfg).float().eval()
    ec = {"state_manager": {"max_context": 256, "max_ragged_batch_size": 64, "max_ragged_sequence_count": 4,
                            "memory_config": {"mode": "allocate", "size": 8}}}
    e = build_engine_from_model(m, ec)
    ids = torch.randint(0, 128, (9, ))
    got = e.put([0], [ids])[0]
    with torch.no_grad():
        ref = m(ids[None])
    ref = ref[0] if isinstance(ref, tuple) else ref
    ref = ref.logits if hasattr(ref, "logits") else ref
    torch.testing.assert_close(got, ref[0, -1].float(), atol=2e-4, rtol=1e-3)
    ec["quantization"] = {"quantization_mode": "int8"}
    eq = build_engine_from_model(m, ec)
    gq = eq.put([0], [ids])[0]
    assert torch.nn.functional.cosine_similarity(gq, got, dim=0) > 0.99


def test_mixtral_routed_path(monkeypatch):
    from deepspeed_b200.inference.v2.model_implementations import ragged_transformer as RT
    monkeypatch.setattr(RT, "MOE_DENSE_MAX_TOKENS", 0)
    test_family_parity("mixtral")
