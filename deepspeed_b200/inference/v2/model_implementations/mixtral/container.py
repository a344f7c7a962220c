"""Checkpoint-name mapping for Mixtral 8x7B-style sparse MoE (top-2 of 8 gated experts) (reference ``model_implementations/mixtral/container.py``)."""
from ..common_parameters import *  # noqa: F401,F403
from .. import param_maps as P
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

    PARAM_MAPPING = {**P.split_qkv("self_attn"), **P.attn_out("self_attn.o_proj"),
                     **P.routed_experts("block_sparse_moe.gate", "block_sparse_moe.experts", "w1", "w3", "w2"),
                     **P.norm("input_layernorm", "attn_norm_gamma"), **P.norm("post_attention_layernorm", "mlp_norm_gamma")}


class MixtralNonTransformerContainer(LayerContainer):
    """Embedding, final norm, LM head."""
    word_emb: EmbeddingParameter
    word_unembed: UnembedParameter
    final_norm: NormParameter

    PARAM_MAPPING = P.embeddings("model.embed_tokens", "model.norm", "lm_head")
