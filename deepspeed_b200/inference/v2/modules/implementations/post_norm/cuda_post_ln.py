"""Post-LayerNorm block (reference ``modules/implementations/post_norm/cuda_post_ln.py``)."""
from typing import Any, Dict

import torch

from ....inference_utils import DtypeEnum, NormTypeEnum
from ....kernels.core_ops import CUDAFPPostLN
from ...configs import DSNormConfig
from ...interfaces import DSPostNormBase, DSPostNormRegistry


@DSPostNormRegistry.register_module
class DSPostLNCUDAModule(DSPostNormBase):

    @staticmethod
    def name() -> str:
        return "cuda_post_ln"

    @staticmethod
    def supports_config(config: DSNormConfig) -> bool:
        return NormTypeEnum(config.type) == NormTypeEnum.LayerNorm and len({config.residual_dtype, config.input_dtype,
                                                                            config.output_dtype}) == 1

    def __init__(self, config: DSNormConfig, implementation_config: Dict[str, Any] = None) -> None:
        super().__init__(config, implementation_config)
        self.ln = CUDAFPPostLN(config.channels, DtypeEnum(config.residual_dtype).value, config.eps)

    def forward(self, residual, hidden_in, gamma, beta=None) -> torch.Tensor:
        self.ln(residual, residual, hidden_in, gamma, beta)
        return residual
