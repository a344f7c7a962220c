"""Norm module config (reference ``modules/configs/norm_config.py``)."""
from deepspeed_b200.inference.v2.inference_utils import DtypeEnum, NormTypeEnum

from ..ds_module import DSModuleConfig

_HALF = DtypeEnum.fp16


class DSNormConfig(DSModuleConfig):
    """``type``: layer_norm | rms_norm over ``channels``; the three dtypes are the residual stream, the module input and
    the normalised output (the CUDA implementations require them equal)."""
    type: NormTypeEnum
    channels: int
    eps: float = 1e-5
    residual_dtype: DtypeEnum = _HALF
    input_dtype: DtypeEnum = _HALF
    output_dtype: DtypeEnum = _HALF
