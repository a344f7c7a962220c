from .config import LoRAConfig, QuantizationConfig  # noqa: F401
from .optimized_linear import OptimizedLinear, LoRAOptimizedLinear  # noqa: F401
from .quantization import QuantizedParameter, QuantizedLinear  # noqa: F401
from .context_manager import Init, init_lora  # noqa: F401
