"""Checkpoint-name mapping for Qwen2-MoE (routed experts + a gated shared expert) (reference ``model_implementations/qwen_v2_moe/container.py``)."""
from ..common_parameters import *  # noqa: F401,F403
from .. import param_maps as P
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

    PARAM_MAPPING = {**P.split_qkv("self_attn", bias=True), **P.attn_out("self_attn.o_proj"),
                     **P.routed_experts("mlp.gate", "mlp.experts", "gate_proj", "up_proj", "down_proj"),
                     "mlp.shared_expert.gate_proj.weight": "shared_moe_mlp_1.gate_params",
                     "mlp.shared_expert.up_proj.weight": "shared_moe_mlp_1.up_params",
                     "mlp.shared_expert.down_proj.weight": "shared_moe_mlp_2.params",
                     "mlp.shared_expert_gate.weight": "shared_moe_gate.params",
                     **P.norm("input_layernorm", "attn_norm_gamma"), **P.norm("post_attention_layernorm", "mlp_norm_gamma")}


class Qwen2MoeNonTransformerContainer(LayerContainer):
    """Embedding, final norm, LM head."""
    word_emb: EmbeddingParameter
    word_unembed: UnembedParameter
    final_norm: NormParameter

    PARAM_MAPPING = P.embeddings("model.embed_tokens", "model.norm", "lm_head")
