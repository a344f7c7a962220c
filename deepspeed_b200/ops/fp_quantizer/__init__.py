from .quantize import FP_Quantize, Quantizer  # noqa: F401
from .fp8_gemm import matmul_fp8  # noqa: F401
