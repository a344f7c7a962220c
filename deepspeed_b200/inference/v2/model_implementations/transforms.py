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

    # ---- architecture description in the vocabulary of the module layer (reference ``DSTransformerModelBase`` properties)
    @property
    def max_sequence_length(self):
        return self.spec.max_positions

    @property
    def positional_embedding_type(self):
        from ..modules.configs import PositionalEmbeddingType
        return PositionalEmbeddingType.rotate_half if self.spec.positional == "rope" else PositionalEmbeddingType.none

    @property
    def positional_embedding_config(self):
        from ..modules.configs import RotateHalfConfig
        if self.spec.positional != "rope":
            return None
        return RotateHalfConfig(theta_base=float(self.spec.rope_theta), rotate_dim=self.spec.rotary_dim)

    @property
    def gated_mlp(self):
        return bool(self.spec.gated_mlp)

    @property
    def mlp_activation_fn(self):
        from ..inference_utils import ActivationType as A
        table = {("silu", True): A.SiGLU, ("gelu", True): A.GEGLU, ("gelu_new", True): A.GEGLU, ("relu", True): A.ReGLU,
                 ("silu", False): A.SILU, ("gelu", False): A.GELU, ("gelu_new", False): A.GELU, ("relu", False): A.RELU}
        return table[(self.spec.act, bool(self.spec.gated_mlp))]

    @property
    def norm_type(self):
        from ..inference_utils import NormTypeEnum
        return NormTypeEnum.RMSNorm if self.spec.norm == "rms" else NormTypeEnum.LayerNorm

    @property
    def n_heads_q_local(self):
        return self.spec.heads // self.tp_size

    @property
    def n_heads_kv_local(self):
        return max(1, self.spec.kv_heads // self.tp_size)

    @property
    def n_top_k(self):
        return self.spec.top_k

    @property
    def normalize_expert_scores(self):
        return bool(self.spec.norm_topk)

    @property
    def model_config(self):
        return self.spec

    @property
    def engine_config(self):
        return getattr(self, "_engine_config", None)

    # ---- module-layer builders: the same architecture expressed as DSModules (``modules/heuristics``) ------------------
    # ``RaggedTransformer.forward`` calls the fused kernels directly; these builders give tools and custom pipelines the
    # reference's per-layer objects, configured from the same ArchSpec.
    def _dt(self):
        from ..inference_utils import DtypeEnum
        return DtypeEnum(self.dtype)

    def make_embedding_layer(self):
        from ..modules import heuristics as H
        from ..modules.configs import DSEmbeddingsConfig
        self.embed = H.instantiate_embed(DSEmbeddingsConfig(max_tokens=self._max_tokens(), residual_dtype=self._dt(),
                                                            embedding_dim=self.spec.hidden,
                                                            positional_embedding=self.spec.positional == "learned",
                                                            positional_offset=self.spec.pos_offset), self.engine_config)
        return self.embed

    def make_unembedding_layer(self):
        from ..modules import heuristics as H
        from ..modules.configs import DSUnembedConfig
        self.unembed = H.instantiate_unembed(DSUnembedConfig(max_tokens=self._max_tokens(), dtype=self._dt(),
                                                             norm_type=self.norm_type if self.spec.final_norm else None,
                                                             model_dim=self.spec.hidden, vocab_size=self.spec.vocab_size),
                                             self.engine_config)
        return self.unembed

    def _max_tokens(self):
        sm = getattr(self.engine_config, "state_manager", None)
        return int(getattr(sm, "max_ragged_batch_size", 768) or 768)

    def _linear(self, cin, cout, act=None):
        from ..inference_utils import ActivationType
        from ..modules import heuristics as H
        from ..modules.configs import DSLinearConfig
        return H.instantiate_linear(DSLinearConfig(max_tokens=self._max_tokens(), in_channels=cin, out_channels=cout,
                                                   activation=act if act is not None else ActivationType.IDENTITY,
                                                   input_dtype=self._dt(), output_dtype=self._dt()), self.engine_config)

    def make_qkv_layer(self):
        out = self.head_size * (self.n_heads_q_local + 2 * self.n_heads_kv_local)
        self.qkv = self._linear(self.spec.hidden, out)
        return self.qkv

    def make_attn_out_layer(self):
        self.attn_out = self._linear(self.head_size * self.n_heads_q_local, self.spec.hidden)
        return self.attn_out

    def make_mlp_1_layer(self):
        self.mlp_1 = self._linear(self.spec.hidden, self.spec.intermediate // self.tp_size, self.mlp_activation_fn)
        return self.mlp_1

    def make_mlp_2_layer(self):
        self.mlp_2 = self._linear(self.spec.intermediate // self.tp_size, self.spec.hidden)
        return self.mlp_2

    def make_norm_layer(self):
        from ..modules import heuristics as H
        from ..modules.configs import DSNormConfig
        self.norm = H.instantiate_pre_norm(DSNormConfig(max_tokens=self._max_tokens(), type=self.norm_type,
                                                        channels=self.spec.hidden, eps=self.spec.norm_eps,
                                                        residual_dtype=self._dt(), input_dtype=self._dt(),
                                                        output_dtype=self._dt()), self.engine_config)
        return self.norm

    def make_attn_layer(self):
        from ..modules import heuristics as H
        from ..modules.configs import DSSelfAttentionConfig
        self.attn = H.instantiate_attention(DSSelfAttentionConfig(
            max_tokens=self._max_tokens(), n_heads_q=self.n_heads_q_local, n_heads_kv=self.n_heads_kv_local,
            head_size=self.head_size, scale_factor=self.head_size**-0.5, input_dtype=self._dt(), output_dtype=self._dt(),
            positional_embedding_type=self.positional_embedding_type,
            positional_embedding_config=self.positional_embedding_config), self.engine_config)
        return self.attn

    def make_moe_layer(self):
        from ..modules import heuristics as H
        from ..modules.configs import DSMoEConfig
        self.moe = H.instantiate_moe(DSMoEConfig(max_tokens=self._max_tokens(), model_dim=self.spec.hidden,
                                                 intermediate_features=self.spec.intermediate // self.tp_size,
                                                 n_experts=self.spec.num_experts, top_k=self.spec.top_k,
                                                 activation=self.mlp_activation_fn, input_dtype=self._dt(),
                                                 output_dtype=self._dt(), normalize_scores=bool(self.spec.norm_topk)),
                                     self.engine_config)
        return self.moe

    def prepare_batch(self, wrapped_batch) -> None:
        """Per-forward device-side preparation (attention atoms in the reference); the paged attention kernel consumes the
        ragged metadata directly, so this only finalises the batch if the caller has not."""
        if hasattr(wrapped_batch, "finalize") and not getattr(wrapped_batch, "_finalized", True):
            wrapped_batch.finalize()

    def set_parameters(self, transformer_containers, non_transformer_container, flattened_param_buffer=None,
                       flattened_param_metadata=None):
        """Keep the declarative containers the weights came from (serialisation walks them)."""
        self._transformer_params, self._non_transformer = transformer_containers, non_transformer_container
        self._flattened_param_buffer, self._flattened_param_metadata = flattened_param_buffer, flattened_param_metadata

    @property
    def flattened_params(self):
        return getattr(self, "_flattened_param_buffer", None)

    @property
    def flattened_param_metadata(self):
        return getattr(self, "_flattened_param_metadata", None)

    @property
    def config(self):
        return self.spec

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
