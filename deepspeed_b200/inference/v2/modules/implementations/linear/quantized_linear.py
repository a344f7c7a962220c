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


def float_quantize(x: torch.Tensor, exp_bits: int, man_bits: int) -> torch.Tensor:
    """Round ``x`` (fp32) to the nearest value representable with ``exp_bits`` / ``man_bits`` (IEEE-style bias, subnormals,
    no inf/nan codes: the top exponent is an ordinary binade; ties to even) and saturate at the format's maximum."""
    bias = 2**(exp_bits - 1) - 1
    e_min, e_max = 1 - bias, 2**exp_bits - 1 - bias
    max_val = 2.0**e_max * (2.0 - 2.0**-man_bits)
    ax = x.abs().clamp_max(max_val)
    e = torch.floor(torch.log2(ax.clamp_min(2.0**(e_min - man_bits - 1)))).clamp(e_min, e_max)
    step = torch.exp2(e - man_bits)
    q = torch.round(ax / step) * step  # torch.round is round-half-to-even
    return torch.copysign(q.clamp_max(max_val), x)


def fp_quantize(input: torch.Tensor, num_bits: int = 6, exp_bits: int = 3, min_value: torch.Tensor = None,
                max_value: torch.Tensor = None, group_size: int = -1):
    """Per-output-channel FP6 (e3m2) fake quantisation (reference ``quantized_linear.py:25``, which needs ``qtorch``): returns
    ``(values on the fp6 grid stored as fp16, fp16 scales)`` with ``dequantised = values * scales``."""
    assert (min_value is None) == (max_value is None)
    assert input.dtype == torch.float16
    if not (num_bits == 6 and exp_bits == 3):
        raise NotImplementedError("only FP6 e3m2 is supported")
    if group_size != -1:
        raise NotImplementedError("only per-channel quantisation (group_size=-1) is supported")
    q_range = 28.0  # largest e3m2 magnitude
    shape = input.shape
    rows = input.float().reshape(-1, shape[-1])
    peak = rows.abs().amax(dim=-1, keepdim=True) if min_value is None else torch.max(min_value.abs(), max_value).float().reshape(-1, 1)
    scales = peak / q_range
    scales[scales == 0] = 1
    fake = float_quantize(rows / scales, exp_bits, num_bits - exp_bits - 1)
    return fake.reshape(shape).contiguous().to(torch.float16), scales.to(torch.float16)
