"""Checkpoint-name mapping for Phi-2 (parallel residual, partial rotary, biased linears) (reference ``model_implementations/phi/container.py``)."""
from ..common_parameters import *  # noqa: F401,F403
from ..layer_container_base import LayerContainer


class PhiTransformerContainer(LayerContainer):
    """One decoder layer (names relative to ``model.layers.<i>.``)."""
    qkv_w: UnfusedQKVParameter
    qkv_b: UnfusedQKVParameter
    attn_out_w: AttentionOutputParameter
    attn_out_b: AttentionOutputParameter
    mlp_1_w: MLP1Parameter
    mlp_1_b: MLP1Parameter
    mlp_2_w: MLP2Parameter
    mlp_2_b: MLP2Parameter
    ln_gamma: NormParameter
    ln_beta: NormParameter

    PARAM_MAPPING = {
        "self_attn.q_proj.weight": "qkv_w.q_params",
        "self_attn.k_proj.weight": "qkv_w.k_params",
        "self_attn.v_proj.weight": "qkv_w.v_params",
        "self_attn.q_proj.bias": "qkv_b.q_params",
        "self_attn.k_proj.bias": "qkv_b.k_params",
        "self_attn.v_proj.bias": "qkv_b.v_params",
        "self_attn.dense.weight": "attn_out_w.params",
        "self_attn.dense.bias": "attn_out_b.params",
        "mlp.fc1.weight": "mlp_1_w.params",
        "mlp.fc1.bias": "mlp_1_b.params",
        "mlp.fc2.weight": "mlp_2_w.params",
        "mlp.fc2.bias": "mlp_2_b.params",
        "input_layernorm.weight": "ln_gamma.params",
        "input_layernorm.bias": "ln_beta.params",
    }


class PhiNonTransformerContainer(LayerContainer):
    """Embedding, final norm, LM head."""
    word_emb: EmbeddingParameter
    word_unembed_w: UnembedParameter
    word_unembed_b: UnembedParameter
    final_norm_gamma: NormParameter
    final_norm_beta: NormParameter

    PARAM_MAPPING = {
        "model.embed_tokens.weight": "word_emb.params",
        "model.final_layernorm.weight": "final_norm_gamma.params",
        "model.final_layernorm.bias": "final_norm_beta.params",
        "lm_head.weight": "word_unembed_w.params",
        "lm_head.bias": "word_unembed_b.params",
    }
