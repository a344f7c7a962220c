"""Checkpoint-name mapping for Qwen-1 (fused c_attn with biases, w1/w2 gated MLP) (reference ``model_implementations/qwen/container.py``)."""
from ..common_parameters import *  # noqa: F401,F403
from .. import param_maps as P
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

    PARAM_MAPPING = {**P.fused_qkv("attn.c_attn", bias=True), **P.attn_out("attn.c_proj"), **P.gated_mlp("mlp.w2", "mlp.w1", "mlp.c_proj"),
                     **P.norm("ln_1", "attn_norm_gamma"), **P.norm("ln_2", "mlp_norm_gamma")}


class QwenNonTransformerContainer(LayerContainer):
    """Embedding, final norm, LM head."""
    word_emb: EmbeddingParameter
    word_unembed: UnembedParameter
    final_norm: NormParameter

    PARAM_MAPPING = P.embeddings("transformer.wte", "transformer.ln_f", "lm_head")
