"""Floating-point linear with fused bias + activation (reference ``modules/implementations/linear/blas_fp_linear.py``)."""
from typing import Any, Dict

import torch

from deepspeed_b200.ops.kernels import transformer_ops as T

from ....inference_utils import ActivationType, DtypeEnum, is_gated
from ....kernels.core_ops import BlasLibLinear
from ...configs import DSLinearConfig
from ...interfaces import DSLinearBase, DSLinearRegistry

_ACT = {ActivationType.GELU: "gelu", ActivationType.RELU: "relu", ActivationType.SILU: "silu", ActivationType.IDENTITY: None,
        ActivationType.GEGLU: "gelu", ActivationType.ReGLU: "relu", ActivationType.SiGLU: "silu"}


@DSLinearRegistry.register_module
class BlasFPLinear(DSLinearBase):

    @staticmethod
    def name() -> str:
        return "blas_fp_linear"

    @staticmethod
    def supports_config(config: DSLinearConfig) -> bool:
        return config.input_dtype == config.output_dtype and config.quantization_mode is None

    def __init__(self, config: DSLinearConfig, implementation_config: Dict[str, Any] = None) -> None:
        super().__init__(config, implementation_config)
        self.gemm = BlasLibLinear(DtypeEnum(config.input_dtype).value)
        self.act = ActivationType(config.activation)
        self._out = None

    @property
    def output(self) -> torch.Tensor:
        return self._out

    def forward(self, hidden_states, w, b=None) -> torch.Tensor:
        rows = w.shape[0]
        y = torch.empty(*hidden_states.shape[:-1], rows, dtype=hidden_states.dtype, device=hidden_states.device)
        self.gemm(y, hidden_states, w)
        if is_gated(self.act):
            y = T.gated_act((y if b is None else y + b).contiguous(), act=_ACT[self.act])
        elif b is not None or _ACT[self.act] is not None:
            y = T.bias_act(y, b, act=_ACT[self.act])
        self._out = y
        return y
