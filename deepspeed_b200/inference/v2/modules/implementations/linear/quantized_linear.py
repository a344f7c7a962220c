"""Weight-only-quantised linears (reference ``modules/implementations/linear/quantized_linear.py``): fp6 (FP6-LLM role),
int8 and int4 weights with per-group scales; activations stay fp16 / bf16."""
from typing import Any, Dict

import torch

from deepspeed_b200.inference.quantization.layers import maybe_quantized_linear, quantize_weight
from deepspeed_b200.ops.kernels import transformer_ops as T

from ....inference_utils import ActivationType, is_gated
from ...configs import DSLinearConfig
from ...interfaces import DSLinearBase, DSLinearRegistry
from .blas_fp_linear import _ACT

_MODES = {"wf6af16": "fp6", "fp6": "fp6", "int8": "int8", "w8a16": "int8", "int4": "int4", "w4a16": "int4", "fp8": "fp8"}


@DSLinearRegistry.register_module
class QuantizedWf6Af16Linear(DSLinearBase):

    @staticmethod
    def name() -> str:
        return "quantized_wf6af16_linear"

    @staticmethod
    def supports_config(config: DSLinearConfig) -> bool:
        return config.input_dtype == config.output_dtype and str(config.quantization_mode).lower() in _MODES

    def __init__(self, config: DSLinearConfig, implementation_config: Dict[str, Any] = None) -> None:
        super().__init__(config, implementation_config)
        self.mode = _MODES[str(config.quantization_mode).lower()]
        self.group_size = int((implementation_config or {}).get("group_size", 128))
        self.act = ActivationType(config.activation)
        self._out = None

    @property
    def output(self) -> torch.Tensor:
        return self._out

    def transform_param(self, param: torch.Tensor):
        """2-D weights are quantised once at load time; biases pass through."""
        if param.dim() != 2 or param.shape[1] % self.group_size != 0:
            return param
        return quantize_weight(param, self.mode, self.group_size)

    def forward(self, hidden_states, w, b=None) -> torch.Tensor:
        y = maybe_quantized_linear(hidden_states, w, None)
        if is_gated(self.act):
            y = T.gated_act((y if b is None else y + b).contiguous(), act=_ACT[self.act])
        elif b is not None or _ACT[self.act] is not None:
            y = T.bias_act(y, b, act=_ACT[self.act])
        self._out = y
        return y
