"""Checkpoint-name mapping for Falcon (parallel attention/MLP, multi-query or grouped fused QKV) (reference ``model_implementations/falcon/container.py``)."""
from ..common_parameters import *  # noqa: F401,F403
from .. import param_maps as P
from ..layer_container_base import LayerContainer


class FalconTransformerContainer(LayerContainer):
    """One decoder layer (names relative to ``transformer.h.<i>.``)."""
    qkv_w: FusedQKVParameter
    attn_out_w: AttentionOutputParameter
    mlp_1_w: MLP1Parameter
    mlp_2_w: MLP2Parameter
    ln_attn_gamma: NormParameter
    ln_attn_beta: NormParameter

    PARAM_MAPPING = {**P.fused_qkv("self_attention.query_key_value"), **P.attn_out("self_attention.dense"),
                     **P.plain_mlp("mlp.dense_h_to_4h", "mlp.dense_4h_to_h"), **P.norm("input_layernorm", "ln_attn_gamma", "ln_attn_beta")}


class FalconNonTransformerContainer(LayerContainer):
    """Embedding, final norm, LM head."""
    word_emb: EmbeddingParameter
    word_unembed: UnembedParameter
    final_norm_w: NormParameter
    final_norm_b: NormParameter

    PARAM_MAPPING = P.embeddings("transformer.word_embeddings", "transformer.ln_f", "lm_head", final_norm_bias=True)


class FalconNewArchTransformerContainer(LayerContainer):
    """Decoder layer of the ``new_decoder_architecture`` checkpoints (Falcon-40B/180B): grouped ``[kv_group: q.. k v]`` fused
    QKV and separate layer norms for the attention and MLP branches (names relative to ``transformer.h.<i>.``)."""
    qkv_w: GQAMegatronQKVParameter
    attn_out_w: AttentionOutputParameter
    mlp_1_w: MLP1Parameter
    mlp_2_w: MLP2Parameter
    ln_attn_gamma: NormParameter
    ln_attn_beta: NormParameter
    ln_mlp_gamma: NormParameter
    ln_mlp_beta: NormParameter

    PARAM_MAPPING = {**P.fused_qkv("self_attention.query_key_value"), **P.attn_out("self_attention.dense"),
                     **P.plain_mlp("mlp.dense_h_to_4h", "mlp.dense_4h_to_h"), **P.norm("ln_attn", "ln_attn_gamma", "ln_attn_beta"),
                     **P.norm("ln_mlp", "ln_mlp_gamma", "ln_mlp_beta")}
