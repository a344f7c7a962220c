"""Norm module config (reference ``modules/configs/norm_config.py``)."""
from deepspeed_b200.inference.v2.inference_utils import DtypeEnum, NormTypeEnum

from ..ds_module import DSModuleConfig


class DSNormConfig(DSModuleConfig):
    type: NormTypeEnum
    channels: int
    residual_dtype: DtypeEnum = DtypeEnum.fp16
    input_dtype: DtypeEnum = DtypeEnum.fp16
    output_dtype: DtypeEnum = DtypeEnum.fp16
    eps: float = 1e-5
