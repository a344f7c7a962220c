"""Checkpoint-name mapping for Phi-2 (parallel residual, partial rotary, biased linears) (reference ``model_implementations/phi/container.py``)."""
from ..common_parameters import *  # noqa: F401,F403
from .. import param_maps as P
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

    PARAM_MAPPING = {**P.split_qkv("self_attn", bias=True), **P.attn_out("self_attn.dense", bias=True), **P.plain_mlp("mlp.fc1", "mlp.fc2", bias=True),
                     **P.norm("input_layernorm", "ln_gamma", "ln_beta")}


class PhiNonTransformerContainer(LayerContainer):
    """Embedding, final norm, LM head."""
    word_emb: EmbeddingParameter
    word_unembed_w: UnembedParameter
    word_unembed_b: UnembedParameter
    final_norm_gamma: NormParameter
    final_norm_beta: NormParameter

    PARAM_MAPPING = {"model.embed_tokens.weight": "word_emb.params", **P.norm("model.final_layernorm", "final_norm_gamma", "final_norm_beta"),
                     "lm_head.weight": "word_unembed_w.params", "lm_head.bias": "word_unembed_b.params"}
