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
