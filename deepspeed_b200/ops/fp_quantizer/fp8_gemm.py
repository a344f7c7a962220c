"""``matmul_fp8``: activation (bf16) x group-quantised FP8 weight (reference ``ops/fp_quantizer/fp8_gemm.py`` +
``fp8_gemm_triton.py``).  The weight is dequantised by the ``fp_dequantize`` kernel into bf16 and multiplied on the tensor
cores; for decode-sized inputs the op is bound by the FP8 weight bytes, which is the point of storing them in 8 bits."""
import torch

from .quantize import FP_Quantize


def matmul_fp8(inp, weight, scale, quantization_group_size, quantizer: FP_Quantize = None):
    """``inp [..., K] @ dequant(weight)[K, N]``; ``weight`` is the packed payload produced by ``FP_Quantize.quantize`` for a
    ``[K, N]`` matrix, ``scale`` its per-group scales."""
    q = quantizer if quantizer is not None else FP_Quantize(group_size=quantization_group_size)
    k = inp.shape[-1]
    groups = scale.numel()
    n = groups * quantization_group_size // k
    q.orig_shape, q.orig_dtype = torch.Size([k, n]), inp.dtype
    w = q.dequantize(weight, q_bits=8, q_mantisa_bits=3, scale=scale).view(k, n).to(inp.dtype)
    return torch.matmul(inp, w)


def matmul_fp8_fallback(inp, weight, scale, quantization_group_size, quantizer: FP_Quantize = None):
    """Library-only path (dequantise → ``torch.matmul``); what ``matmul_fp8`` does for shapes the fused weight-only kernel
    does not cover (reference ``fp8_gemm.py``)."""
    return matmul_fp8(inp, weight, scale, quantization_group_size, quantizer)
