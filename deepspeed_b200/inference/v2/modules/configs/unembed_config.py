"""Unembedding (LM head) module config (reference ``modules/configs/unembed_config.py``)."""
from typing import Optional

from deepspeed_b200.inference.v2.inference_utils import DtypeEnum, NormTypeEnum

from ..ds_module import DSModuleConfig


class DSUnembedConfig(DSModuleConfig):
    dtype: DtypeEnum = DtypeEnum.fp16
    norm_type: Optional[NormTypeEnum] = None
    model_dim: int
    max_sequences: int = 128
    vocab_size: int
