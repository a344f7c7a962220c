"""Checkpoint-name mapping for Falcon (parallel attention/MLP, multi-query or grouped fused QKV) (reference ``model_implementations/falcon/container.py``)."""
from ..common_parameters import *  # noqa: F401,F403
from ..layer_container_base import LayerContainer


class FalconTransformerContainer(LayerContainer):
    """One decoder layer (names relative to ``transformer.h.<i>.``)."""
    qkv_w: FusedQKVParameter
    attn_out_w: AttentionOutputParameter
    mlp_1_w: MLP1Parameter
    mlp_2_w: MLP2Parameter
    ln_attn_gamma: NormParameter
    ln_attn_beta: NormParameter

    PARAM_MAPPING = {
        "self_attention.query_key_value.weight": "qkv_w.params",
        "self_attention.dense.weight": "attn_out_w.params",
        "mlp.dense_h_to_4h.weight": "mlp_1_w.params",
        "mlp.dense_4h_to_h.weight": "mlp_2_w.params",
        "input_layernorm.weight": "ln_attn_gamma.params",
        "input_layernorm.bias": "ln_attn_beta.params",
    }


class FalconNonTransformerContainer(LayerContainer):
    """Embedding, final norm, LM head."""
    word_emb: EmbeddingParameter
    word_unembed: UnembedParameter
    final_norm_w: NormParameter
    final_norm_b: NormParameter

    PARAM_MAPPING = {
        "transformer.word_embeddings.weight": "word_emb.params",
        "transformer.ln_f.weight": "final_norm_w.params",
        "transformer.ln_f.bias": "final_norm_b.params",
        "lm_head.weight": "word_unembed.params",
    }
