"""Transformer specialisation of the model protocol (reference ``model_implementations/inference_transformer_base.py``).
``RaggedTransformer`` is the implementation; this module pins the property names family code relies on."""
from .inference_model_base import DSInferenceModelBase
from .ragged_transformer import RaggedTransformer


class DSTransformerModelBase(DSInferenceModelBase):
    """Property protocol of a decoder-only transformer."""
    num_layers: int
    model_dim: int
    vocab_size: int
    head_size: int
    n_heads: int
    intermediate_dim: int
    n_heads_kv: int


DSTransformerModelBase.register(RaggedTransformer)


class DSMoETransformerModelBase(DSTransformerModelBase):
    """Adds the routed-expert properties (reference ``inference_transformer_base.py:532``): ``RaggedTransformer`` serves
    dense and MoE families alike, its ``ArchSpec`` carries the expert count / top-k / score normalisation."""
    n_experts: int
    n_top_k: int
    normalize_expert_scores: bool

    @classmethod
    def __subclasshook__(cls, other):
        return NotImplemented

    @staticmethod
    def is_moe(model) -> bool:
        return int(getattr(getattr(model, "spec", None), "num_experts", 0) or 0) > 0


DSMoETransformerModelBase.register(RaggedTransformer)
