"""Module layer of the ragged engine: registries, heuristics and every implementation vs plain torch (host tier)."""
from types import SimpleNamespace

import pytest
import torch

from deepspeed_b200.inference.v2.inference_utils import ActivationType, DtypeEnum, NormTypeEnum, ceil_div, elem_size, is_gated
from deepspeed_b200.inference.v2.modules import heuristics as H
from deepspeed_b200.inference.v2.modules.configs import (DSEmbeddingsConfig, DSLinearConfig, DSMoEConfig, DSNormConfig,
                                                         DSSelfAttentionConfig, DSUnembedConfig, PositionalEmbeddingType,
                                                         RotateHalfConfig)
from deepspeed_b200.inference.v2.modules.interfaces import DSLinearRegistry
from deepspeed_b200.inference.v2.modules.module_registry import ConfigBundle


class _Batch:
    """Two sequences (5 and 3 tokens), block size 4."""

    def __init__(self):
        self._ids = torch.randint(0, 40, (8, ), dtype=torch.int32)
        self._seq = torch.tensor([0] * 5 + [1] * 3, dtype=torch.int32)
        self._pos = torch.tensor(list(range(5)) + list(range(3)), dtype=torch.int32)
        self._bt = torch.tensor([[0, 1], [2, 3]], dtype=torch.int32)

    input_ids = lambda self: self._ids
    seq_of = lambda self: self._seq
    pos_of = lambda self: self._pos
    block_table = lambda self: self._bt
    last_token_index = lambda self: torch.tensor([4, 7], dtype=torch.int32)


def test_enums_and_helpers():
    assert DtypeEnum("bf16").value is torch.bfloat16 and DtypeEnum(torch.float16) is DtypeEnum.fp16
    assert is_gated(ActivationType.SiGLU) and not is_gated(0) and elem_size(torch.bfloat16) == 2 and ceil_div(7, 4) == 2
    with pytest.raises(KeyError):
        DSLinearRegistry.instantiate_config(ConfigBundle(name="nope", config=None))


def test_linear_norm_embed_unembed_modules():
    torch.manual_seed(0)
    b = _Batch()
    fp32 = DtypeEnum.fp32
    lin = H.instantiate_linear(DSLinearConfig(in_channels=16, out_channels=24, activation=ActivationType.SiGLU, input_dtype=fp32,
                                              output_dtype=fp32))
    x, w, bias = torch.randn(8, 16), torch.randn(48, 16), torch.randn(48)
    y = lin(x, w, bias)
    h = x @ w.t() + bias
    assert torch.allclose(y, torch.nn.functional.silu(h[:, :24]) * h[:, 24:], atol=1e-5)
    qlin = H.instantiate_linear(DSLinearConfig(in_channels=128, out_channels=16, input_dtype=DtypeEnum.bf16,
                                               output_dtype=DtypeEnum.bf16, quantization_mode="int8"))
    assert type(qlin).__name__ == "QuantizedWf6Af16Linear"
    wq = torch.randn(16, 128).bfloat16() * 0.1
    qw = qlin.transform_param(wq)
    xq = torch.randn(3, 128).bfloat16()
    assert (qlin(xq, qw).float() - xq.float() @ wq.float().t()).abs().max() < 0.06
    pre = H.instantiate_pre_norm(DSNormConfig(type=NormTypeEnum.RMSNorm, channels=16, residual_dtype=fp32, input_dtype=fp32,
                                              output_dtype=fp32, eps=1e-6))
    res, delta, g = torch.randn(8, 16), torch.randn(8, 16), torch.rand(16)
    want = res + delta
    r2, hid = pre(res.clone(), delta, g)
    assert torch.allclose(r2, want, atol=1e-6)
    assert torch.allclose(hid, want * torch.rsqrt(want.pow(2).mean(-1, keepdim=True) + 1e-6) * g, atol=1e-5)
    r3, hid0 = pre(res.clone(), None, g)
    assert torch.equal(r3, res)
    post = H.instantiate_post_norm(DSNormConfig(type=NormTypeEnum.LayerNorm, channels=16, residual_dtype=fp32, input_dtype=fp32,
                                                output_dtype=fp32))
    out = post(res.clone(), delta, g, torch.zeros(16))
    assert torch.allclose(out, torch.nn.functional.layer_norm(want, (16, ), g, torch.zeros(16), 1e-5), atol=1e-5)
    emb = H.instantiate_embed(DSEmbeddingsConfig(embedding_dim=16, residual_dtype=fp32, positional_embedding=True, positional_offset=2))
    wte, wpe = torch.randn(40, 16), torch.randn(12, 16)
    e = emb(b, wte, wpe)
    assert torch.allclose(e, wte[b.input_ids().long()] + wpe[b.pos_of().long() + 2])
    un = H.instantiate_unembed(DSUnembedConfig(dtype=fp32, norm_type=NormTypeEnum.RMSNorm, model_dim=16, vocab_size=40))
    logits = un(e, wte, b, gamma=g)
    last = e[[4, 7]]
    ref = (last * torch.rsqrt(last.pow(2).mean(-1, keepdim=True) + 1e-5) * g) @ wte.t()
    assert logits.shape == (2, 40) and torch.allclose(logits, ref, atol=1e-4)


def test_attention_and_moe_modules():
    torch.manual_seed(0)
    b = _Batch()
    fp32 = DtypeEnum.fp32
    attn = H.instantiate_attention(DSSelfAttentionConfig(n_heads_q=4, n_heads_kv=2, head_size=16, input_dtype=fp32, output_dtype=fp32,
                                                         positional_embedding_type=PositionalEmbeddingType.rotate_half,
                                                         positional_embedding_config=RotateHalfConfig(theta_base=10000.0)),
                                   SimpleNamespace(state_manager=SimpleNamespace(max_context=16)))
    attn._block = 4
    cache = torch.zeros(4, 4, 2, 2, 16)
    qkv = torch.randn(8, 8 * 16)
    out = attn(qkv.clone(), cache, b)
    assert out.shape == (8, 64) and torch.isfinite(out).all() and cache.abs().sum() > 0
    # first token of each sequence attends only to itself: output = its own V (GQA: 2 q heads per kv head)
    v0 = qkv.view(8, 8, 16)[0, 6:]
    assert torch.allclose(out[0].view(4, 16), v0.repeat_interleave(2, 0), atol=1e-5)
    moe = H.instantiate_moe(DSMoEConfig(model_dim=16, intermediate_features=32, n_experts=4, top_k=2, input_dtype=fp32,
                                        output_dtype=fp32, activation=ActivationType.GELU, normalize_scores=True))
    x = torch.randn(6, 16)
    gw, w1, w2 = torch.randn(4, 16), torch.randn(4, 32, 16) * 0.2, torch.randn(4, 16, 32) * 0.2
    y = moe(x, gw, w1, w2)
    probs = (x @ gw.t()).softmax(-1)
    tv, ti = probs.topk(2, -1)
    tv = tv / tv.sum(-1, keepdim=True)
    ref = torch.zeros_like(x)
    for t in range(6):
        for j in range(2):
            e = int(ti[t, j])
            ref[t] += tv[t, j] * (torch.nn.functional.gelu(x[t] @ w1[e].t()) @ w2[e].t())
    assert torch.allclose(y, ref, atol=1e-4), (y - ref).abs().max()


def test_allocator_parameter_and_checkpoint_engines(tmp_path):
    from deepspeed_b200.inference.v2.allocator import empty_from
    from deepspeed_b200.inference.v2.checkpoint import InMemoryModelEngine
    from deepspeed_b200.inference.v2.inference_parameter import InferenceParameter
    from deepspeed_b200.inference.v2.logging import inference_logger
    buf = torch.zeros(64)
    v = empty_from(buf, (4, 8))
    v.fill_(1.0)
    assert buf[:32].sum() == 32 and buf[32:].sum() == 0
    with pytest.raises(ValueError):
        empty_from(buf, (65, ))
    p = InferenceParameter.initialize(torch.ones(4, dtype=torch.float16), scales=torch.full((2, ), 3.0))
    q = p.to(torch.float32)
    assert q.dtype == torch.float32 and torch.equal(q.scales, torch.full((2, ), 3.0)) and "scales" in q.aux_attrs
    m = torch.nn.Linear(3, 2)
    eng = InMemoryModelEngine(m)
    assert dict(eng.parameters())["weight"] is not None and eng.get("bias").shape == (2, )
    assert inference_logger() is inference_logger()
