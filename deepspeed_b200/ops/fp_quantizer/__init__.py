from .quantize import FP_Quantize, Quantizer  # noqa: F401
