"""Checkpoint-name mapping for Mixtral 8x7B-style sparse MoE (top-2 of 8 gated experts) (reference ``model_implementations/mixtral/container.py``)."""
from ..common_parameters import *  # noqa: F401,F403
from ..layer_container_base import LayerContainer


class MixtralTransformerContainer(LayerContainer):
    """One decoder layer (names relative to ``model.layers.<i>.``)."""
    qkv_w: UnfusedQKVParameter
    attn_out_w: AttentionOutputParameter
    moe_gate: MoEGatingWeightParameter
    moe_mlp_1: UnfusedMoEGatedMLPParameter
    moe_mlp_2: UnfusedMoEMLP2Parameter
    attn_norm_gamma: NormParameter
    mlp_norm_gamma: NormParameter

    PARAM_MAPPING = {
        "self_attn.q_proj.weight": "qkv_w.q_params",
        "self_attn.k_proj.weight": "qkv_w.k_params",
        "self_attn.v_proj.weight": "qkv_w.v_params",
        "self_attn.o_proj.weight": "attn_out_w.params",
        "block_sparse_moe.gate.weight": "moe_gate.params",
        "block_sparse_moe.experts.*.w1.weight": "moe_mlp_1.gating_experts",
        "block_sparse_moe.experts.*.w3.weight": "moe_mlp_1.up_experts",
        "block_sparse_moe.experts.*.w2.weight": "moe_mlp_2.experts",
        "input_layernorm.weight": "attn_norm_gamma.params",
        "post_attention_layernorm.weight": "mlp_norm_gamma.params",
    }


class MixtralNonTransformerContainer(LayerContainer):
    """Embedding, final norm, LM head."""
    word_emb: EmbeddingParameter
    word_unembed: UnembedParameter
    final_norm: NormParameter

    PARAM_MAPPING = {
        "model.embed_tokens.weight": "word_emb.params",
        "model.norm.weight": "final_norm.params",
        "lm_head.weight": "word_unembed.params",
    }
