"""Embedding module config (reference ``modules/configs/embedding_config.py``)."""
from typing import Optional

from deepspeed_b200.inference.v2.inference_utils import DtypeEnum, NormTypeEnum

from ..ds_module import DSModuleConfig


class DSEmbeddingsConfig(DSModuleConfig):
    residual_dtype: DtypeEnum = DtypeEnum.fp16
    embedding_dim: int
    positional_embedding: bool = False
    positional_offset: int = 0  # OPT stores position p at row p + 2
    use_token_type: bool = False
    output_normalization: Optional[NormTypeEnum] = None
