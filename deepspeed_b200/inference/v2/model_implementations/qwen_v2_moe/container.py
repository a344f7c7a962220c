"""Checkpoint-name mapping for Qwen2-MoE (routed experts + a gated shared expert) (reference ``model_implementations/qwen_v2_moe/container.py``)."""
from ..common_parameters import *  # noqa: F401,F403
from ..layer_container_base import LayerContainer


class Qwen2MoeTransformerContainer(LayerContainer):
    """One decoder layer (names relative to ``model.layers.<i>.``)."""
    qkv_w: UnfusedQKVParameter
    qkv_b: UnfusedQKVParameter
    attn_out_w: AttentionOutputParameter
    moe_gate: MoEGatingWeightParameter
    moe_mlp_1: UnfusedMoEGatedMLPParameter
    moe_mlp_2: UnfusedMoEMLP2Parameter
    shared_moe_mlp_1: GatedMLPParameter
    shared_moe_mlp_2: MLP2Parameter
    shared_moe_gate: MoEGatingWeightParameter
    attn_norm_gamma: NormParameter
    mlp_norm_gamma: NormParameter

    PARAM_MAPPING = {
        "self_attn.q_proj.weight": "qkv_w.q_params",
        "self_attn.k_proj.weight": "qkv_w.k_params",
        "self_attn.v_proj.weight": "qkv_w.v_params",
        "self_attn.q_proj.bias": "qkv_b.q_params",
        "self_attn.k_proj.bias": "qkv_b.k_params",
        "self_attn.v_proj.bias": "qkv_b.v_params",
        "self_attn.o_proj.weight": "attn_out_w.params",
        "mlp.gate.weight": "moe_gate.params",
        "mlp.experts.*.gate_proj.weight": "moe_mlp_1.gating_experts",
        "mlp.experts.*.up_proj.weight": "moe_mlp_1.up_experts",
        "mlp.experts.*.down_proj.weight": "moe_mlp_2.experts",
        "mlp.shared_expert.gate_proj.weight": "shared_moe_mlp_1.gate_params",
        "mlp.shared_expert.up_proj.weight": "shared_moe_mlp_1.up_params",
        "mlp.shared_expert.down_proj.weight": "shared_moe_mlp_2.params",
        "mlp.shared_expert_gate.weight": "shared_moe_gate.params",
        "input_layernorm.weight": "attn_norm_gamma.params",
        "post_attention_layernorm.weight": "mlp_norm_gamma.params",
    }


class Qwen2MoeNonTransformerContainer(LayerContainer):
    """Embedding, final norm, LM head."""
    word_emb: EmbeddingParameter
    word_unembed: UnembedParameter
    final_norm: NormParameter

    PARAM_MAPPING = {
        "model.embed_tokens.weight": "word_emb.params",
        "model.norm.weight": "final_norm.params",
        "lm_head.weight": "word_unembed.params",
    }
