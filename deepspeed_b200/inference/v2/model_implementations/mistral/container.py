"""Checkpoint-name mapping for Mistral (llama layout, sliding-window capable) (reference ``model_implementations/mistral/container.py``)."""
from ..common_parameters import *  # noqa: F401,F403
from .. import param_maps as P
from ..layer_container_base import LayerContainer


class MistralTransformerContainer(LayerContainer):
    """One decoder layer (names relative to ``model.layers.<i>.``)."""
    qkv_w: UnfusedQKVParameter
    attn_out_w: AttentionOutputParameter
    mlp_1_w: GatedMLPParameter
    mlp_2_w: MLP2Parameter
    attn_norm_gamma: NormParameter
    mlp_norm_gamma: NormParameter

    PARAM_MAPPING = {**P.split_qkv("self_attn"), **P.attn_out("self_attn.o_proj"), **P.gated_mlp("mlp.gate_proj", "mlp.up_proj", "mlp.down_proj"),
                     **P.norm("input_layernorm", "attn_norm_gamma"), **P.norm("post_attention_layernorm", "mlp_norm_gamma")}


class MistralNonTransformerContainer(LayerContainer):
    """Embedding, final norm, LM head."""
    word_emb: EmbeddingParameter
    word_unembed: UnembedParameter
    final_norm: NormParameter

    PARAM_MAPPING = P.embeddings("model.embed_tokens", "model.norm", "lm_head")
