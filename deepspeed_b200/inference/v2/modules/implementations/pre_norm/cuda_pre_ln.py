"""Pre-LayerNorm block (reference ``modules/implementations/pre_norm/cuda_pre_ln.py``)."""
from typing import Any, Dict

import torch

from ....inference_utils import DtypeEnum, NormTypeEnum
from ....kernels.core_ops import CUDAFPLN, CUDAFPPreLN
from ...configs import DSNormConfig
from ...interfaces import DSPreNormBase, DSPreNormRegistry


@DSPreNormRegistry.register_module
class DSPreLNCUDAModule(DSPreNormBase):

    @staticmethod
    def name() -> str:
        return "cuda_pre_ln"

    @staticmethod
    def supports_config(config: DSNormConfig) -> bool:
        return NormTypeEnum(config.type) == NormTypeEnum.LayerNorm and len({config.residual_dtype, config.input_dtype,
                                                                            config.output_dtype}) == 1

    def __init__(self, config: DSNormConfig, implementation_config: Dict[str, Any] = None) -> None:
        super().__init__(config, implementation_config)
        dt = DtypeEnum(config.residual_dtype).value
        self.ln, self.pre_ln = CUDAFPLN(config.channels, dt, config.eps), CUDAFPPreLN(config.channels, dt, config.eps)

    def forward(self, residual, hidden_in, gamma, beta=None):
        hidden = torch.empty_like(residual)
        if hidden_in is None:
            self.ln(hidden, residual, gamma, beta)
        else:
            self.pre_ln(residual, hidden, residual, hidden_in, gamma, beta)
        return residual, hidden
