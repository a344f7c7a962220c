"""Checkpoint-name mapping for OPT (LayerNorm, learned positions offset by 2, ReLU) (reference ``model_implementations/opt/container.py``)."""
from ..common_parameters import *  # noqa: F401,F403
from ..layer_container_base import LayerContainer


class OPTTransformerContainer(LayerContainer):
    """One decoder layer (names relative to ``model.decoder.layers.<i>.``)."""
    qkv_w: UnfusedQKVParameter
    qkv_b: UnfusedQKVParameter
    attn_out_w: AttentionOutputParameter
    attn_out_b: AttentionOutputParameter
    mlp_1_w: MLP1Parameter
    mlp_1_b: MLP1Parameter
    mlp_2_w: MLP2Parameter
    mlp_2_b: MLP2Parameter
    attn_norm_gamma: NormParameter
    attn_norm_beta: NormParameter
    mlp_norm_gamma: NormParameter
    mlp_norm_beta: NormParameter

    PARAM_MAPPING = {
        "self_attn.q_proj.weight": "qkv_w.q_params",
        "self_attn.k_proj.weight": "qkv_w.k_params",
        "self_attn.v_proj.weight": "qkv_w.v_params",
        "self_attn.q_proj.bias": "qkv_b.q_params",
        "self_attn.k_proj.bias": "qkv_b.k_params",
        "self_attn.v_proj.bias": "qkv_b.v_params",
        "self_attn.out_proj.weight": "attn_out_w.params",
        "self_attn.out_proj.bias": "attn_out_b.params",
        "fc1.weight": "mlp_1_w.params",
        "fc1.bias": "mlp_1_b.params",
        "fc2.weight": "mlp_2_w.params",
        "fc2.bias": "mlp_2_b.params",
        "self_attn_layer_norm.weight": "attn_norm_gamma.params",
        "self_attn_layer_norm.bias": "attn_norm_beta.params",
        "final_layer_norm.weight": "mlp_norm_gamma.params",
        "final_layer_norm.bias": "mlp_norm_beta.params",
    }


class OPTNonTransformerContainer(LayerContainer):
    """Embedding, final norm, LM head."""
    word_emb: EmbeddingParameter
    word_emb_pos: EmbeddingParameter
    word_unembed: UnembedParameter
    final_norm_w: NormParameter
    final_norm_b: NormParameter

    PARAM_MAPPING = {
        "model.decoder.embed_tokens.weight": ["word_emb.params", "word_unembed.params"],
        "model.decoder.embed_positions.weight": "word_emb_pos.params",
        "model.decoder.final_layer_norm.weight": "final_norm_w.params",
        "model.decoder.final_layer_norm.bias": "final_norm_b.params",
    }
