from .quantizer import (Quantizer, ds_quantizer, quantize, dequantize, swizzle_quant, quantized_reduction,  # noqa: F401
                        loco_quantized_reduction, fake_quantize)
