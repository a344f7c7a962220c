"""Layer-wise kernel injection (policy -> container -> fused layer) reproduces the Hugging Face model for every
registered family, on tiny random-weight models (fp32, host tier)."""
import copy
from types import SimpleNamespace

import pytest
import torch

transformers = pytest.importorskip("transformers")
from transformers import AutoConfig, AutoModel, AutoModelForCausalLM  # noqa: E402

from deepspeed_b200.module_inject.containers.base import InjectedLayer  # noqa: E402
from deepspeed_b200.module_inject.replace_module import replace_transformer_layer  # noqa: E402


def _cfg(**kw):
    return SimpleNamespace(replace_with_kernel_inject=True, dtype=torch.float32, max_out_tokens=64,
                           tensor_parallel=SimpleNamespace(tp_size=1),
                           quant=SimpleNamespace(enabled=False, weight=SimpleNamespace(post_init_quant=None)), **kw)


CAUSAL = {
    "gpt2": dict(vocab_size=100, n_embd=32, n_layer=2, n_head=4, n_positions=64),
    "opt": dict(vocab_size=100, hidden_size=32, num_hidden_layers=2, num_attention_heads=4, ffn_dim=64, max_position_embeddings=64,
                word_embed_proj_dim=32),
    "llama": dict(vocab_size=100, hidden_size=32, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2,
                  intermediate_size=64, max_position_embeddings=64),
    "mistral": dict(vocab_size=100, hidden_size=32, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2,
                    intermediate_size=64, max_position_embeddings=64, sliding_window=None),
    "gptj": dict(vocab_size=100, n_embd=32, n_layer=2, n_head=4, n_positions=64, rotary_dim=4),
    "gpt_neo": dict(vocab_size=100, hidden_size=32, num_layers=2, num_heads=4, max_position_embeddings=64,
                    attention_types=[[["global", "local"], 1]], window_size=4),
    "gpt_neox": dict(vocab_size=100, hidden_size=32, num_hidden_layers=2, num_attention_heads=4, intermediate_size=64,
                     max_position_embeddings=64, rotary_pct=0.5),
    "bloom": dict(vocab_size=100, hidden_size=32, n_layer=2, n_head=4),
}


@pytest.mark.parametrize("family", sorted(CAUSAL))
def test_causal_family_matches_hf(family):
    torch.manual_seed(0)
    cfg = AutoConfig.for_model(family, **CAUSAL[family])
    cfg._attn_implementation = "eager"
    model = AutoModelForCausalLM.from_config(cfg).eval()
    ids = torch.randint(0, 100, (2, 10))
    with torch.no_grad():
        ref = model(ids).logits
    inj = replace_transformer_layer(None, copy.deepcopy(model), config=_cfg())
    n = sum(isinstance(m, InjectedLayer) for m in inj.modules())
    assert n == 2, f"{family}: {n} layers injected"
    with torch.no_grad():
        out = inj(ids).logits
    assert (out - ref).abs().max() < 2e-4, f"{family}: max err {(out - ref).abs().max()}"


@pytest.mark.parametrize("family", ["gpt2", "llama"])
def test_incremental_decoding_through_hf_generate(family):
    """``generate`` with HF's cache plumbing: the fused layers own the KV cache, the HF cache only tracks length."""
    torch.manual_seed(0)
    cfg = AutoConfig.for_model(family, **CAUSAL[family])
    cfg._attn_implementation = "eager"
    model = AutoModelForCausalLM.from_config(cfg).eval()
    inj = replace_transformer_layer(None, copy.deepcopy(model), config=_cfg())
    ids = torch.randint(0, 100, (2, 6))
    kw = dict(max_new_tokens=8, do_sample=False, pad_token_id=0)
    with torch.no_grad():
        ref = model.generate(ids, **kw)
        out = inj.generate(ids, **kw)
    assert torch.equal(ref, out)


ENCODERS = {
    "bert": dict(vocab_size=100, hidden_size=32, num_hidden_layers=2, num_attention_heads=4, intermediate_size=64,
                 max_position_embeddings=64),
    "distilbert": dict(vocab_size=100, dim=32, n_layers=2, n_heads=4, hidden_dim=64, max_position_embeddings=64),
    "roberta": dict(vocab_size=100, hidden_size=32, num_hidden_layers=2, num_attention_heads=4, intermediate_size=64,
                    max_position_embeddings=66),
}


@pytest.mark.parametrize("family", sorted(ENCODERS))
def test_encoder_family_matches_hf(family):
    torch.manual_seed(0)
    cfg = AutoConfig.for_model(family, **ENCODERS[family])
    cfg._attn_implementation = "eager"
    model = AutoModel.from_config(cfg).eval()
    ids = torch.randint(2, 100, (2, 9))
    mask = torch.ones(2, 9, dtype=torch.long)
    mask[1, 6:] = 0
    with torch.no_grad():
        ref = model(ids, attention_mask=mask).last_hidden_state
    inj = replace_transformer_layer(None, copy.deepcopy(model), config=_cfg())
    assert sum(isinstance(m, InjectedLayer) for m in inj.modules()) == 2
    with torch.no_grad():
        out = inj(ids, attention_mask=mask).last_hidden_state
    keep = mask.bool()
    assert (out[keep] - ref[keep]).abs().max() < 2e-4


def test_clip_text_tower_and_checkpoint_loading():
    torch.manual_seed(0)
    from transformers import CLIPTextConfig, CLIPTextModel
    cfg = CLIPTextConfig(vocab_size=100, hidden_size=32, intermediate_size=64, num_hidden_layers=2, num_attention_heads=4,
                         max_position_embeddings=16)
    cfg._attn_implementation = "eager"
    model = CLIPTextModel(cfg).eval()
    ids = torch.randint(0, 99, (2, 8))
    with torch.no_grad():
        ref = model(ids).last_hidden_state
    inj = replace_transformer_layer(None, copy.deepcopy(model), config=_cfg())
    with torch.no_grad():
        assert (inj(ids).last_hidden_state - ref).abs().max() < 2e-4
    # checkpoint-driven loading: scramble the injected weights, reload them from the original state dict
    from deepspeed_b200.module_inject.load_checkpoint import load_model_with_checkpoint
    torch.manual_seed(1)
    lcfg = AutoConfig.for_model("llama", **CAUSAL["llama"])
    lcfg._attn_implementation = "eager"
    lm = AutoModelForCausalLM.from_config(lcfg).eval()
    sd = {k: v.clone() for k, v in lm.state_dict().items()}
    inj = replace_transformer_layer(None, copy.deepcopy(lm), config=_cfg())
    with torch.no_grad():
        for p in inj.parameters():
            p.normal_()
    n = load_model_with_checkpoint(inj, sd)
    assert n >= 4
    x = torch.randint(0, 100, (1, 7))
    with torch.no_grad():
        assert (inj(x).logits - lm(x).logits).abs().max() < 2e-4


def test_structural_policies_megatron_and_llama2_reference_layout():
    """Families that cannot be imported here are recognised by structure."""
    from torch import nn
    from deepspeed_b200.module_inject.containers import LLAMA2LayerPolicy, MegatronLayerPolicy
    from deepspeed_b200.module_inject.replace_policy import policy_for

    class Attn(nn.Module):

        def __init__(self):
            super().__init__()
            self.query_key_value, self.dense = nn.Linear(16, 48), nn.Linear(16, 16)
            self.num_attention_heads = 4

    class MLP(nn.Module):

        def __init__(self):
            super().__init__()
            self.dense_h_to_4h, self.dense_4h_to_h = nn.Linear(16, 64), nn.Linear(64, 16)

    class ParallelTransformerLayer(nn.Module):

        def __init__(self):
            super().__init__()
            self.input_layernorm, self.post_attention_layernorm = nn.LayerNorm(16), nn.LayerNorm(16)
            self.self_attention, self.mlp = Attn(), MLP()

    layer = ParallelTransformerLayer()
    assert policy_for(layer) is MegatronLayerPolicy
    pol = MegatronLayerPolicy(layer)
    assert pol.get_hidden_heads()[:2] == (16, 4) and pol.attention()[0].shape == (48, 16)

    class A2(nn.Module):

        def __init__(self):
            super().__init__()
            self.wq, self.wk, self.wv, self.wo = (nn.Linear(16, 16, bias=False) for _ in range(4))
            self.n_heads = 4

    class FF(nn.Module):

        def __init__(self):
            super().__init__()
            self.w1, self.w2, self.w3 = nn.Linear(16, 40, bias=False), nn.Linear(40, 16, bias=False), nn.Linear(16, 40, bias=False)

    class TransformerBlock(nn.Module):

        def __init__(self):
            super().__init__()
            self.attention, self.feed_forward = A2(), FF()
            self.attention_norm, self.ffn_norm = nn.LayerNorm(16), nn.LayerNorm(16)

    blk = TransformerBlock()
    assert policy_for(blk) is LLAMA2LayerPolicy
    assert LLAMA2LayerPolicy(blk).mlp()[0].shape == (80, 16)


def test_tp_shard_and_fused_qkv_helpers():
    from deepspeed_b200.module_inject import tp_shard
    from deepspeed_b200.module_inject.fusedqkv_utils import prepare_tp_fused_qkvw, shard_chunk_mlp
    tp_shard.set_num_kv_heads(6)
    assert tp_shard.get_shard_size_list(6 * 8, 4) == [16, 16, 8, 8]  # whole heads, remainder to the first ranks
    tp_shard.set_num_kv_heads(None)
    tp_shard.set_tp_grain_size(64)
    assert tp_shard.get_shard_size_list(256, 3, "mlp") == [128, 64, 64]
    tp_shard.set_tp_grain_size(1)
    w = torch.arange(24.).reshape(24, 1)  # glm layout: q rows 0-7, k 8-15, v 16-23
    assert prepare_tp_fused_qkvw("glmtype", w, 2, 1).flatten().tolist() == [4, 5, 6, 7, 12, 13, 14, 15, 20, 21, 22, 23]
    assert prepare_tp_fused_qkvw("bloomtype", w, 2, 0).flatten().tolist() == list(map(float, range(12)))
    gw, _ = shard_chunk_mlp(torch.arange(8.).reshape(8, 1), None, 1, 2)
    assert gw.flatten().tolist() == [2, 3, 6, 7]


def test_init_inference_kernel_inject_encoder():
    """``init_inference(replace_with_kernel_inject=True)`` on a model the ragged engine has no entry for (BERT)."""
    import deepspeed_b200 as ds
    from tests.common import run_distributed
    run_distributed(_bert_engine, 1)


def _bert_engine():
    import deepspeed_b200 as ds
    torch.manual_seed(0)
    cfg = AutoConfig.for_model("bert", **ENCODERS["bert"])
    cfg._attn_implementation = "eager"
    model = AutoModel.from_config(cfg).eval()
    ids = torch.randint(2, 100, (2, 9))
    with torch.no_grad():
        ref = model(ids).last_hidden_state
    eng = ds.init_inference(copy.deepcopy(model), dtype=torch.float32, replace_with_kernel_inject=True)
    assert sum(isinstance(m, InjectedLayer) for m in eng.module.modules()) == 2
    with torch.no_grad():
        out = eng(ids).last_hidden_state
    assert (out - ref).abs().max() < 2e-4


def _tp_inject(family):
    """tp=2 layer-wise injection: every rank holds half the heads / MLP columns; the all-reduced result equals the HF model."""
    import torch.distributed as td
    from deepspeed_b200.utils import groups
    torch.manual_seed(0)
    cfg = AutoConfig.for_model(family, **CAUSAL[family])
    cfg._attn_implementation = "eager"
    model = AutoModelForCausalLM.from_config(cfg).eval()
    ids = torch.randint(0, 100, (2, 9))
    with torch.no_grad():
        ref = model(ids).logits
    icfg = _cfg()
    icfg.tensor_parallel = SimpleNamespace(tp_size=2)
    inj = replace_transformer_layer(None, copy.deepcopy(model), config=icfg)
    layers = [m for m in inj.modules() if isinstance(m, InjectedLayer)]
    assert len(layers) == 2
    f = layers[0].fused
    full = [m for m in model.modules() if type(m).__name__.endswith(("DecoderLayer", "Block"))][0]
    assert f.attn_qkvw.shape[0] < sum(p.shape[0] for n, p in full.named_parameters() if any(k in n for k in ("q_proj.weight", "k_proj.weight",
                                      "v_proj.weight", "c_attn.weight"))) or family == "gpt2"
    with torch.no_grad():
        out = inj(ids).logits
    assert (out - ref).abs().max() < 3e-4, (out - ref).abs().max()
    # checkpoint-driven loading applies the same slicing
    from deepspeed_b200.module_inject.load_checkpoint import load_model_with_checkpoint
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    with torch.no_grad():
        for l in layers:
            for p in l.fused.parameters():
                p.normal_()
    load_model_with_checkpoint(inj, sd, mp_group=groups.get_tensor_model_parallel_group(), mp_size=2)
    with torch.no_grad():
        assert (inj(ids).logits - ref).abs().max() < 3e-4


@pytest.mark.parametrize("family", ["llama", "gpt2", "opt"])
def test_layerwise_injection_tensor_parallel(family):
    from tests.common import run_distributed
    run_distributed(_tp_inject, 2, (family, ))


def test_module_quantize_and_training_inject():
    from transformers.models.bert.modeling_bert import BertLayer
    from deepspeed_b200.module_inject.inject import module_inject
    from deepspeed_b200.module_inject.module_quantize import quantize_transformer_layer
    torch.manual_seed(0)
    cfg = AutoConfig.for_model("bert", **ENCODERS["bert"])
    cfg._attn_implementation = "eager"
    model = AutoModel.from_config(cfg).eval()
    q = quantize_transformer_layer(BertLayer, copy.deepcopy(model))
    lin = q.encoder.layer[0].intermediate.dense
    assert lin.weight.dtype == torch.int8 and hasattr(lin.weight, "scale")
    ref_w = model.encoder.layer[0].intermediate.dense.weight
    assert (lin.weight.float() * lin.weight.scale - ref_w).abs().max() <= lin.weight.scale * 0.51
    # training-time injection: fused training layer with the HF weights copied in
    tr = module_inject(BertLayer, copy.deepcopy(model), cfg, micro_batch_size=2, max_seq_length=16, seed=1, preln=False, fp16=False)
    from deepspeed_b200.ops.transformer import DeepSpeedTransformerLayer
    new = tr.encoder.layer[0]
    assert isinstance(new, DeepSpeedTransformerLayer)
    a = model.encoder.layer[0].attention
    assert torch.equal(new.attn_qkvw[:32], a.self.query.weight) and torch.equal(new.output_w, model.encoder.layer[0].output.dense.weight)


@pytest.mark.parametrize("family", ["gpt2", "bloom", "gpt_neox", "opt"])
def test_checkpoint_loading_per_family(family):
    """load_model_with_checkpoint understands each family's checkpoint layout (Conv1D transposes, per-head fused QKV, ...)."""
    from deepspeed_b200.module_inject.load_checkpoint import load_model_with_checkpoint
    torch.manual_seed(0)
    cfg = AutoConfig.for_model(family, **CAUSAL[family])
    cfg._attn_implementation = "eager"
    model = AutoModelForCausalLM.from_config(cfg).eval()
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    inj = replace_transformer_layer(None, copy.deepcopy(model), config=_cfg())
    with torch.no_grad():
        for m in inj.modules():
            if isinstance(m, InjectedLayer):
                for p in m.fused.parameters():
                    p.normal_()
    assert load_model_with_checkpoint(inj, sd) >= 2
    ids = torch.randint(0, 100, (2, 7))
    with torch.no_grad():
        assert (inj(ids).logits - model(ids).logits).abs().max() < 3e-4


def test_alibi_helpers_and_fused_qkv_layouts():
    from deepspeed_b200.module_inject.auto_tp_model_utils import build_bloom_alibi_tensor, get_alibi_mask
    from deepspeed_b200.module_inject.fusedqkv_utils import fused_type_of, prepare_tp_fused_qkvw, require_tp_fused_qkvw
    from transformers.models.bloom.modeling_bloom import build_alibi_tensor
    mask = torch.ones(2, 6, dtype=torch.long)
    mask[1, :2] = 0
    ours = build_bloom_alibi_tensor(mask, 4, torch.float32)
    ref = build_alibi_tensor(mask, 4, torch.float32)
    assert ours.shape == ref.shape and torch.allclose(ours, ref, atol=1e-6)
    from types import SimpleNamespace
    m = get_alibi_mask(SimpleNamespace(n_head=4), torch.zeros(1), 5)
    assert m.shape == (4, 5, 5) and torch.isinf(m[0, 0, 1]) and m[0, 1, 0] < 0 and m[0, 2, 2] == 0
    assert fused_type_of("BloomBlock") == "bloomtype" and require_tp_fused_qkvw("h.0.self_attention.query_key_value", 2)
    assert not require_tp_fused_qkvw("h.0.mlp.fc1", 2)
    # codegen: 4 groups of [q | v | k]; each rank takes its slice of every block of every group
    w = torch.arange(48.).reshape(48, 1)
    a, b = (prepare_tp_fused_qkvw("codegentype", w, 2, r).flatten() for r in range(2))
    assert a.numel() == 24 and sorted(torch.cat([a, b]).tolist()) == list(map(float, range(48)))
    assert a[:2].tolist() == [0.0, 1.0] and b[:2].tolist() == [2.0, 3.0]
