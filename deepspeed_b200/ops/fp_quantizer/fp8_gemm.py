"""``matmul_fp8``: activation (bf16) x group-quantised FP8 weight (reference ``ops/fp_quantizer/fp8_gemm.py`` +
``fp8_gemm_triton.py``).

Two implementations, picked by the weight's storage:

* the weight was quantised as ``[N, K]`` rows with groups along K (``QuantizedWeight`` of
  ``inference/quantization/layers.py``) -> the fused kernels stream the FP8 bytes and dequantise on chip: ``wq_gemm.cu``
  (mma.sync, <= 32 rows) or ``wq_tc_gemm.cu`` (tcgen05, dequantise-in-shared-memory, up to ~1k rows);
* the reference calling convention -- packed payload of a ``[K, N]`` matrix + per-group scales -> :func:`matmul_fp8_fallback`
  (the ``fp_dequantize`` kernel into bf16, then a tensor-core GEMM).
"""
import torch

from .quantize import FP_Quantize


def matmul_fp8_fallback(inp, weight, scale, quantization_group_size, quantizer: FP_Quantize = None):
    """Dequantise -> GEMM: ``inp [..., K] @ dequant(weight)[K, N]`` for the packed payload ``FP_Quantize.quantize``
    produced from a ``[K, N]`` matrix (``scale``: its per-group scales)."""
    q = quantizer if quantizer is not None else FP_Quantize(group_size=quantization_group_size)
    k = inp.shape[-1]
    groups = scale.numel()
    n = groups * quantization_group_size // k
    q.orig_shape, q.orig_dtype = torch.Size([k, n]), inp.dtype
    w = q.dequantize(weight, q_bits=8, q_mantisa_bits=3, scale=scale).view(k, n).to(inp.dtype)
    return torch.matmul(inp, w)


def matmul_fp8(inp, weight, scale=None, quantization_group_size=None, quantizer: FP_Quantize = None):
    """FP8-weight GEMM.  ``weight`` may be a ``QuantizedWeight`` (``mode == "fp8"``, ``[N, K]`` layout: fused on-chip
    dequantisation, nothing but FP8 bytes leave HBM) or the reference's packed ``[K, N]`` payload with ``scale``."""
    from deepspeed_b200.inference.quantization.layers import QuantizedWeight, maybe_quantized_linear
    if isinstance(weight, QuantizedWeight):
        return maybe_quantized_linear(inp, weight)
    return matmul_fp8_fallback(inp, weight, scale, quantization_group_size, quantizer)
