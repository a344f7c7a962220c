"""Checkpoint-name mapping for Qwen-1 (fused c_attn with biases, w1/w2 gated MLP) (reference ``model_implementations/qwen/container.py``)."""
from ..common_parameters import *  # noqa: F401,F403
from ..layer_container_base import LayerContainer


class QwenTransformerContainer(LayerContainer):
    """One decoder layer (names relative to ``transformer.h.<i>.``)."""
    qkv_w: FusedQKVParameter
    qkv_b: FusedQKVParameter
    attn_out_w: AttentionOutputParameter
    mlp_1_w: GatedMLPParameter
    mlp_2_w: MLP2Parameter
    attn_norm_gamma: NormParameter
    mlp_norm_gamma: NormParameter

    PARAM_MAPPING = {
        "attn.c_attn.weight": "qkv_w.params",
        "attn.c_attn.bias": "qkv_b.params",
        "attn.c_proj.weight": "attn_out_w.params",
        "mlp.w2.weight": "mlp_1_w.gate_params",
        "mlp.w1.weight": "mlp_1_w.up_params",
        "mlp.c_proj.weight": "mlp_2_w.params",
        "ln_1.weight": "attn_norm_gamma.params",
        "ln_2.weight": "mlp_norm_gamma.params",
    }


class QwenNonTransformerContainer(LayerContainer):
    """Embedding, final norm, LM head."""
    word_emb: EmbeddingParameter
    word_unembed: UnembedParameter
    final_norm: NormParameter

    PARAM_MAPPING = {
        "transformer.wte.weight": "word_emb.params",
        "transformer.ln_f.weight": "final_norm.params",
        "lm_head.weight": "word_unembed.params",
    }
