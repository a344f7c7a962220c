"""Sharding arithmetic, declarative containers, per-family policies and the flat-model serialisation of the ragged engine."""
import pytest
import torch

from deepspeed_b200.inference.v2.model_implementations.sharding import (ShardingType, get_local_heads, get_shard_endpoints,
                                                                        shard_attn_out_param, shard_mlp_1_param,
                                                                        shard_mlp_2_param, shard_param, shard_qkv_param,
                                                                        shard_unembed_param)


def test_shard_endpoints_cover_the_dimension():
    for size, n, g in ((256, 3, 32), (128, 4, 32), (96, 5, 32), (7, 2, 1)):
        spans = [get_shard_endpoints(size, r, n, g) for r in range(n)]
        assert spans[0][0] == 0 and spans[-1][1] == size
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert all((e - s) % g == 0 for s, e in spans)


def test_qkv_and_attention_output_sharding_reassembles():
    torch.manual_seed(0)
    d, hq, hkv, H = 8, 8, 2, 32
    w = torch.randn((hq + 2 * hkv) * d, H)
    # GQA with as many kv heads as ranks
    parts = [shard_qkv_param(w, r, 2, d, hq, hkv) for r in range(2)]
    assert all(p.shape == ((4 + 2) * d, H) for p in parts)
    assert torch.equal(parts[1][:4 * d], w[4 * d:8 * d]) and torch.equal(parts[1][4 * d:5 * d], w[(hq + 1) * d:(hq + 2) * d])
    # fewer kv heads than ranks: kv head replicated
    parts4 = [shard_qkv_param(w, r, 4, d, hq, hkv) for r in range(4)]
    assert get_local_heads(3, 4, hq, hkv) == (2, 1)
    assert torch.equal(parts4[0][2 * d:3 * d], parts4[1][2 * d:3 * d]) and torch.equal(parts4[2][2 * d:3 * d], w[(hq + 1) * d:(hq + 2) * d])
    # MHA uneven split (6 heads over 4 ranks: 2,2,1,1)
    w6 = torch.randn(3 * 6 * d, H)
    sizes = [shard_qkv_param(w6, r, 4, d).shape[0] // (3 * d) for r in range(4)]
    assert sizes == [2, 2, 1, 1]
    o = torch.randn(H, hq * d)
    cols = [shard_attn_out_param(o, r, 2, d, hq, hkv) for r in range(2)]
    assert torch.equal(torch.cat(cols, 1), o)
    assert shard_attn_out_param(torch.ones(H), 1, 2, d, hq, hkv) is None  # bias only on rank 0


def test_mlp_and_unembed_sharding_matches_dense_math():
    torch.manual_seed(0)
    x = torch.randn(3, 64)
    gate_up = torch.randn(2 * 128, 64)
    down = torch.randn(64, 128)
    dense = (torch.nn.functional.silu(x @ gate_up[:128].t()) * (x @ gate_up[128:].t())) @ down.t()
    acc = torch.zeros_like(dense)
    for r in range(4):
        w1 = shard_mlp_1_param(gate_up, r, 4, gated=True)
        w2 = shard_mlp_2_param(down, r, 4)
        h = x @ w1.t()
        g, u = h.chunk(2, -1)
        acc += (torch.nn.functional.silu(g) * u) @ w2.t()
    assert torch.allclose(acc, dense, atol=1e-4)
    vocab = torch.randn(50, 64)
    assert torch.equal(torch.cat([shard_unembed_param(vocab, r, 3) for r in range(3)], 0), vocab)
    moe = torch.randn(4, 2 * 64, 32)
    assert shard_mlp_1_param(moe, 1, 2, gated=True, is_moe=True).shape == (4, 64, 32)
    assert shard_param(torch.ones(8), ShardingType.INNER_DIMENSION, 1, 2) is None


def _tiny_llama():
    from transformers import AutoConfig, AutoModelForCausalLM
    cfg = AutoConfig.for_model("llama", vocab_size=96, hidden_size=64, num_hidden_layers=2, num_attention_heads=4,
                               num_key_value_heads=2, intermediate_size=128, max_position_embeddings=64)
    torch.manual_seed(0)
    return cfg, AutoModelForCausalLM.from_config(cfg).eval()


def test_containers_collect_and_transform_a_checkpoint():
    pytest.importorskip("transformers")
    from deepspeed_b200.inference.v2.checkpoint import InMemoryModelEngine
    from deepspeed_b200.inference.v2.model_implementations.inference_policy_base import policy_for
    cfg, hf = _tiny_llama()
    pol = policy_for("llama")(cfg, InMemoryModelEngine(hf))
    model = pol.instantiate_model(None)
    assert type(model).__name__ == "Llama2InferenceModel" and model.num_layers == 2 and model.n_heads_kv == 2
    cmap = pol.build_container_map(model)
    for name, t in InMemoryModelEngine(hf).parameters():
        assert cmap.map_param(name, t), name
    cmap.validate()
    layer0 = cmap.transformer_params[0]
    sd = hf.state_dict()
    want_qkv = torch.cat([sd[f"model.layers.0.self_attn.{n}_proj.weight"] for n in "qkv"], 0).to(model.dtype)
    assert torch.equal(layer0.qkv_w, want_qkv)
    want_mlp1 = torch.cat([sd["model.layers.0.mlp.gate_proj.weight"], sd["model.layers.0.mlp.up_proj.weight"]], 0).to(model.dtype)
    assert torch.equal(layer0.mlp_1_w, want_mlp1)
    assert cmap.non_transformer_params.word_emb.shape == (96, 64)
    # an incomplete container is reported
    cmap2 = pol.build_container_map(model)
    cmap2.map_param("model.embed_tokens.weight", sd["model.embed_tokens.weight"])
    with pytest.raises(RuntimeError):
        cmap2.validate()


def test_policy_builds_a_serving_model_and_flat_roundtrip(tmp_path):
    pytest.importorskip("transformers")
    from deepspeed_b200.inference.v2.checkpoint import InMemoryModelEngine
    from deepspeed_b200.inference.v2.model_implementations.flat_model_helpers import (flatten_inference_model,
                                                                                     restore_inference_model)
    from deepspeed_b200.inference.v2.model_implementations.inference_policy_base import policy_for
    cfg, hf = _tiny_llama()
    model = policy_for("llama")(cfg, InMemoryModelEngine(hf)).build_model(None)
    buf, meta = flatten_inference_model(model, str(tmp_path / "flat"))
    assert buf.dtype == torch.uint8 and all(m["offset"] % 256 == 0 for m in meta.values()) and (tmp_path / "flat.bin").exists()
    before = {n: t.detach().clone() for n, t in model.flat_tensors().items()}
    assert len(before) > 10, "model exposes no weights to serialise"
    with torch.no_grad():
        for t in model.flat_tensors().values():
            t.zero_()
    restore_inference_model(model, str(tmp_path / "flat"))
    for n, t in model.flat_tensors().items():
        assert torch.equal(t, before[n]), n


def test_moe_container_list_dependencies():
    from types import SimpleNamespace
    from deepspeed_b200.inference.v2.model_implementations.arch import ArchSpec
    from deepspeed_b200.inference.v2.model_implementations.mixtral import MixtralTransformerContainer
    from deepspeed_b200.inference.v2.model_implementations.transforms import ContainerTransformsMixin

    class M(ContainerTransformsMixin):
        spec = ArchSpec("mixtral", 32, 16, 1, 2, 2, 8, 32, num_experts=3, top_k=2)
        tp_size, tp_rank, dtype = 1, 0, torch.float32

    c = MixtralTransformerContainer(M())
    for e in range(3):
        assert c.set_dependency(f"block_sparse_moe.experts.{e}.w1.weight", torch.full((32, 16), float(e)))
        c.set_dependency(f"block_sparse_moe.experts.{e}.w3.weight", torch.full((32, 16), 10.0 + e))
        c.set_dependency(f"block_sparse_moe.experts.{e}.w2.weight", torch.zeros(16, 32))
    assert not c.is_initialized
    c.set_dependency("block_sparse_moe.gate.weight", torch.zeros(3, 16))
    for n in "qkv":
        c.set_dependency(f"self_attn.{n}_proj.weight", torch.zeros(16, 16))
    c.set_dependency("self_attn.o_proj.weight", torch.zeros(16, 16))
    c.set_dependency("input_layernorm.weight", torch.ones(16))
    c.set_dependency("post_attention_layernorm.weight", torch.ones(16))
    assert c.is_initialized and c.moe_mlp_1.shape == (3, 64, 16)
    assert float(c.moe_mlp_1[2, 0, 0]) == 2.0 and float(c.moe_mlp_1[2, 32, 0]) == 12.0
    assert not c.set_dependency("unknown.weight", torch.zeros(1))


def test_parameter_and_container_metaclasses():
    import types
    import torch
    from deepspeed_b200.inference.v2.model_implementations import parameter_base as P
    from deepspeed_b200.inference.v2.model_implementations.layer_container_base import LayerContainer, LayerMetaclass
    from deepspeed_b200.inference.v2.model_implementations.inference_policy_base import POLICIES_BY_NAME, PolicyMeta, InferenceV2Policy

    class Fused(P.ParameterBase):
        a: torch.Tensor
        b: torch.Tensor
        experts = P.ParametrizedList("n_experts")

        def finalize(self):
            return torch.cat([self.a, self.b] + list(self.experts))

    assert isinstance(Fused, P.ParameterMetaclass) and Fused.tensor_dependencies == ("a", "b")
    assert set(Fused.list_dependencies) == {"experts"} and Fused.n_dependencies == 3 and isinstance(Fused.a, property)
    model = types.SimpleNamespace(n_experts=2)
    f = Fused(model)
    f.a, f.b = torch.zeros(1), torch.ones(1)
    f.experts[0] = torch.full((1, ), 2.0)
    assert f.result is None
    with pytest.raises(ValueError):
        f.experts = []
    f.experts[1] = torch.full((1, ), 3.0)
    assert f.result.tolist() == [0.0, 1.0, 2.0, 3.0]

    class Sub(Fused):
        c: torch.Tensor

    assert Sub.tensor_dependencies == ("a", "b", "c")

    class C(LayerContainer):
        w: Fused
        PARAM_MAPPING = {"x.a": "w.a", "x.b": "w.b", "e.*.w": "w.experts"}

    assert isinstance(C, LayerMetaclass) and list(C.annotation_attrs) == ["w"] and len(C._compiled_rules) == 3
    c = C(model)
    for name, t in (("x.a", torch.zeros(1)), ("x.b", torch.ones(1)), ("e.0.w", torch.ones(1)), ("e.1.w", torch.ones(1))):
        assert c.set_dependency(name, t)
    assert c.is_initialized and c.w.numel() == 4 and not c.set_dependency("nope", torch.zeros(1))
    assert isinstance(InferenceV2Policy, PolicyMeta) and "Llama2Policy" in POLICIES_BY_NAME


def test_falcon_new_arch_container_and_misc_v2_helpers():
    import types
    import torch
    from deepspeed_b200.inference.v2.model_implementations.falcon import FalconNewArchTransformerContainer, FalconPolicy
    from deepspeed_b200.inference.v2.model_implementations.common_parameters.mlp_parameters import FusedGatedMLPParameter
    from deepspeed_b200.inference.v2.model_implementations.inference_transformer_base import DSMoETransformerModelBase
    from deepspeed_b200.inference.v2.modules.implementations.linear.quantized_linear import fp_quantize
    from deepspeed_b200.inference.v2.ragged.ragged_wrapper import to_padded
    from deepspeed_b200.inference.v2.kernels.ragged_ops.blocked_flash.blocked_flash import get_kv_block_size, get_q_block_size
    cfg = {"model_type": "falcon", "vocab_size": 64, "hidden_size": 32, "num_hidden_layers": 1, "num_attention_heads": 4,
           "num_kv_heads": 2, "new_decoder_architecture": True, "parallel_attn": True, "bias": False}
    pol = FalconPolicy(cfg)
    model = pol.instantiate_model(None)
    cmap = pol.build_container_map(model)
    layer = cmap.transformer_params[0] if not callable(cmap.transformer_params) else list(cmap.transformer_params())[0]
    assert isinstance(layer, FalconNewArchTransformerContainer)
    for name, shape in (("self_attention.query_key_value.weight", (2 * (2 + 2) * 8, 32)), ("self_attention.dense.weight", (32, 32)),
                        ("mlp.dense_h_to_4h.weight", (128, 32)), ("mlp.dense_4h_to_h.weight", (32, 128)), ("ln_attn.weight", (32, )),
                        ("ln_attn.bias", (32, )), ("ln_mlp.weight", (32, )), ("ln_mlp.bias", (32, ))):
        assert layer.set_dependency(name, torch.randn(*shape)), name
    assert layer.is_initialized
    fake = types.SimpleNamespace(transform_mlp_1_param=lambda t: t)
    p = FusedGatedMLPParameter(fake)
    p.params = torch.arange(8.0).reshape(4, 2)
    assert torch.equal(p.result, torch.arange(8.0).reshape(4, 2))
    assert DSMoETransformerModelBase.is_moe(types.SimpleNamespace(spec=types.SimpleNamespace(num_experts=8)))
    q, s = fp_quantize(torch.randn(4, 32).half())
    assert q.dtype == torch.float16 and s.shape == (4, 1) and float(q.abs().max()) <= 28.0
    assert to_padded(1) == 64 and to_padded(513) == 640 and get_q_block_size(128) == 128 and get_kv_block_size(256) == 64


def test_model_module_layer_builders_and_arch_properties():
    import torch
    from deepspeed_b200.inference.v2.inference_utils import ActivationType, NormTypeEnum
    from deepspeed_b200.inference.v2.model_implementations.llama_v2.model import Llama2InferenceModel
    from deepspeed_b200.inference.v2.model_implementations.mixtral.model import MixtralInferenceModel
    from deepspeed_b200.inference.v2.modules.configs import PositionalEmbeddingType
    cfg = {"model_type": "llama", "vocab_size": 64, "hidden_size": 64, "num_hidden_layers": 1, "num_attention_heads": 4,
           "num_key_value_heads": 2, "intermediate_size": 64, "max_position_embeddings": 128, "rope_theta": 5000.0}
    m = Llama2InferenceModel.from_hf_config(cfg, dtype=torch.float32, device="cpu")
    assert m.mlp_activation_fn == ActivationType.SiGLU and m.norm_type == NormTypeEnum.RMSNorm and m.gated_mlp
    assert m.positional_embedding_type == PositionalEmbeddingType.rotate_half and m.positional_embedding_config.theta_base == 5000.0
    assert (m.n_heads_q_local, m.n_heads_kv_local, m.max_sequence_length) == (4, 2, 128)
    qkv, out, m1, m2, norm = m.make_qkv_layer(), m.make_attn_out_layer(), m.make_mlp_1_layer(), m.make_mlp_2_layer(), m.make_norm_layer()
    x = torch.randn(5, 64)
    w_qkv = qkv.transform_param(torch.randn(16 * (4 + 4), 64))
    assert qkv(x, w_qkv).shape == (5, 128)
    w1 = m1.transform_param(torch.randn(128, 64))
    h = m1(x, w1)
    assert h.shape == (5, 64)  # gated: 2x rows in, intermediate out
    assert m2(h, m2.transform_param(torch.randn(64, 64))).shape == (5, 64)
    assert m.make_attn_layer() is m.attn and m.make_embedding_layer() is m.embed and m.make_unembedding_layer() is m.unembed
    moe_cfg = {"model_type": "mixtral", "vocab_size": 64, "hidden_size": 64, "num_hidden_layers": 1, "num_attention_heads": 4,
               "num_key_value_heads": 2, "intermediate_size": 64, "num_local_experts": 4, "num_experts_per_tok": 2}
    mm = MixtralInferenceModel.from_hf_config(moe_cfg, dtype=torch.float32, device="cpu")
    assert (mm.n_experts, mm.n_top_k) == (4, 2) and mm.make_moe_layer() is mm.moe
