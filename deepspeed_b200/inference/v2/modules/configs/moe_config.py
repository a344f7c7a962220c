"""MoE module config (reference ``modules/configs/moe_config.py``)."""
from deepspeed_b200.inference.v2.inference_utils import ActivationType, DtypeEnum

from ..ds_module import DSModuleConfig


class DSMoEConfig(DSModuleConfig):
    model_dim: int
    intermediate_features: int
    n_experts: int
    top_k: int = 1
    input_dtype: DtypeEnum = DtypeEnum.fp16
    output_dtype: DtypeEnum = DtypeEnum.fp16
    activation: ActivationType = ActivationType.IDENTITY
    normalize_scores: bool = False
