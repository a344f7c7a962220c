"""GEMM backend switch.

``matmul_nt(a, b)`` = ``a @ b.T`` and ``matmul_nn(a, b)`` = ``a @ b`` for 2-D bf16/fp16/fp32 operands.
Backends: ``cublas`` (torch.matmul -> cuBLASLt; the library baseline) and ``sm100`` (the hand-written
tcgen05/TMEM/TMA kernel in ``csrc/cuda/gemm_sm100.cu``).  ``DSB200_GEMM=sm100|cublas|auto``; ``auto``
uses sm100 when the shape is supported (M, N multiples of 128; K multiple of 64; bf16) and the native
kernel passed its self-check on this device, else cuBLAS.
"""
import os

import torch

_backend = os.environ.get("DSB200_GEMM", "auto").lower()
_sm100_ok = None


def set_backend(name: str):
    global _backend
    assert name in ("auto", "cublas", "sm100")
    _backend = name


def get_backend():
    return _backend


def _sm100_usable(a, b, nt):
    global _sm100_ok
    if _backend == "cublas" or not a.is_cuda or a.dtype != torch.bfloat16 or b.dtype != torch.bfloat16:
        return False
    if _sm100_ok is None:
        try:
            from deepspeed_b200.ops.kernels import gemm_sm100
            _sm100_ok = gemm_sm100.self_check()
        except Exception:
            _sm100_ok = False
    if not _sm100_ok:
        if _backend == "sm100":
            raise RuntimeError("DSB200_GEMM=sm100 requested but the native tcgen05 GEMM is unavailable")
        return False
    from deepspeed_b200.ops.kernels import gemm_sm100
    return gemm_sm100.supports(a, b, nt)


def matmul_nt(a, b):
    """a [M, K] @ b[N, K]^T."""
    if _sm100_usable(a, b, True):
        from deepspeed_b200.ops.kernels import gemm_sm100
        return gemm_sm100.matmul_nt(a, b)
    return torch.matmul(a, b.t())


def matmul_nn(a, b):
    """a [M, K] @ b[K, N]."""
    if _sm100_usable(a, b, False):
        from deepspeed_b200.ops.kernels import gemm_sm100
        return gemm_sm100.matmul_nn(a, b)
    return torch.matmul(a, b)
