"""Linear module config (reference ``modules/configs/linear_config.py``)."""
from typing import Optional

from deepspeed_b200.inference.v2.inference_utils import ActivationType, DtypeEnum

from ..ds_module import DSModuleConfig


class DSLinearConfig(DSModuleConfig):
    in_channels: int
    out_channels: int  # for gated activations: the width AFTER gating (the weight has 2x rows)
    activation: ActivationType = ActivationType.IDENTITY
    input_dtype: DtypeEnum = DtypeEnum.fp16
    output_dtype: DtypeEnum = DtypeEnum.fp16
    quantization_mode: Optional[str] = None
