"""Parameter-transform hooks used by the declarative containers: shard for this TP rank, cast to the activation dtype,
optionally quantise (reference: the ``transform_*_param`` methods of ``inference_transformer_base.DSTransformerModelBase``)."""
import torch

from .sharding import (shard_attn_out_param, shard_embedding_param, shard_mlp_1_param, shard_mlp_2_param, shard_qkv_param,
                       shard_unembed_param)


class ContainerTransformsMixin:
    """Expects ``self.spec`` (ArchSpec), ``self.tp_size``, ``self.tp_rank``, ``self.dtype``."""

    # ---- facts the common parameters ask for ------------------------------------------------------------------------
    @property
    def num_layers(self):
        return self.spec.layers

    @property
    def model_dim(self):
        return self.spec.hidden

    @property
    def vocab_size(self):
        return self.spec.vocab_size

    @property
    def head_size(self):
        return self.spec.head_dim

    @property
    def n_heads(self):
        return self.spec.heads

    n_heads_q = n_heads

    @property
    def n_heads_kv(self):
        return self.spec.kv_heads

    @property
    def intermediate_dim(self):
        return self.spec.intermediate

    @property
    def n_experts(self):
        return self.spec.num_experts

    @property
    def activation_dtype(self):
        return self.dtype

    def _cast(self, t):
        return None if t is None else t.to(self.dtype).contiguous()

    # ---- transforms -----------------------------------------------------------------------------------------------------
    def transform_embedding_param(self, p):
        return self._cast(p)  # embeddings are replicated here (the logits all-gather happens on the vocab split)

    def transform_unembed_param(self, p):
        return self._cast(shard_unembed_param(p, self.tp_rank, self.tp_size))

    def transform_qkv_param(self, p):
        return self._cast(shard_qkv_param(p, self.tp_rank, self.tp_size, self.head_size, self.n_heads, self.n_heads_kv))

    def transform_attn_out_param(self, p):
        return self._cast(shard_attn_out_param(p, self.tp_rank, self.tp_size, self.head_size, self.n_heads, self.n_heads_kv))

    def transform_mlp_1_param(self, p):
        gated = self.spec.gated_mlp and p.shape[0] == 2 * self.spec.intermediate
        return self._cast(shard_mlp_1_param(p, self.tp_rank, self.tp_size, gated=gated))

    def transform_mlp_2_param(self, p):
        return self._cast(shard_mlp_2_param(p, self.tp_rank, self.tp_size))

    def transform_moe_gate_param(self, p):
        return self._cast(p)

    def transform_moe_mlp_1_param(self, p):
        return self._cast(shard_mlp_1_param(p, self.tp_rank, self.tp_size, gated=self.spec.gated_mlp, is_moe=True))

    def transform_moe_mlp_2_param(self, p):
        return self._cast(shard_mlp_2_param(p, self.tp_rank, self.tp_size, is_moe=True))

    def transform_norm_param(self, p):
        return self._cast(p)
