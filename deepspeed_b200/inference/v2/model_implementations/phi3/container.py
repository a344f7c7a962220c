"""Checkpoint-name mapping for Phi-3 (fused qkv_proj and gate_up_proj) (reference ``model_implementations/phi3/container.py``)."""
from ..common_parameters import *  # noqa: F401,F403
from .. import param_maps as P
from ..layer_container_base import LayerContainer


class Phi3TransformerContainer(LayerContainer):
    """One decoder layer (names relative to ``model.layers.<i>.``)."""
    qkv_w: FusedQKVParameter
    attn_out_w: AttentionOutputParameter
    mlp_1_w: MLP1Parameter
    mlp_2_w: MLP2Parameter
    attn_norm_gamma: NormParameter
    mlp_norm_gamma: NormParameter

    PARAM_MAPPING = {**P.fused_qkv("self_attn.qkv_proj"), **P.attn_out("self_attn.o_proj"),
                     **P.plain_mlp("mlp.gate_up_proj", "mlp.down_proj"),
                     **P.norm("input_layernorm", "attn_norm_gamma"), **P.norm("post_attention_layernorm", "mlp_norm_gamma")}


class Phi3NonTransformerContainer(LayerContainer):
    """Embedding, final norm, LM head."""
    word_emb: EmbeddingParameter
    word_unembed: UnembedParameter
    final_norm: NormParameter

    PARAM_MAPPING = P.embeddings("model.embed_tokens", "model.norm", "lm_head")
