from .layers import QuantizedWeight, maybe_quantized_linear, quantize_weight  # noqa: F401
from .quantization import _init_group_wise_weight_quantization  # noqa: F401
