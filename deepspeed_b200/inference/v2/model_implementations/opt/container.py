"""Checkpoint-name mapping for OPT (LayerNorm, learned positions offset by 2, ReLU) (reference ``model_implementations/opt/container.py``)."""
from ..common_parameters import *  # noqa: F401,F403
from .. import param_maps as P
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

    PARAM_MAPPING = {**P.split_qkv("self_attn", bias=True), **P.attn_out("self_attn.out_proj", bias=True), **P.plain_mlp("fc1", "fc2", bias=True),
                     **P.norm("self_attn_layer_norm", "attn_norm_gamma", "attn_norm_beta"),
                     **P.norm("final_layer_norm", "mlp_norm_gamma", "mlp_norm_beta")}


class OPTNonTransformerContainer(LayerContainer):
    """Embedding, final norm, LM head."""
    word_emb: EmbeddingParameter
    word_emb_pos: EmbeddingParameter
    word_unembed: UnembedParameter
    final_norm_w: NormParameter
    final_norm_b: NormParameter

    PARAM_MAPPING = {**P.embeddings("model.decoder.embed_tokens", "model.decoder.final_layer_norm", final_norm_bias=True, tie=True),
                     "model.decoder.embed_positions.weight": "word_emb_pos.params"}
