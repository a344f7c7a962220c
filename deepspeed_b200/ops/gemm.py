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


_tuned = {}  # (M, N, K) -> "sm100" | "cublas"


def _pick(a, b):
    """``auto`` mode: time both implementations once per problem shape (CUDA events) and keep the faster
    -- the role of the reference's cuBLAS algorithm sweep (``csrc/includes/gemm_test.h:58``)."""
    key = (a.shape[0], b.shape[0], a.shape[1])
    choice = _tuned.get(key)
    if choice is not None:
        return choice
    if _backend == "sm100" or torch.cuda.is_current_stream_capturing():
        return "sm100"
    from deepspeed_b200.ops.kernels import gemm_sm100
    out = torch.empty(a.shape[0], b.shape[0], dtype=a.dtype, device=a.device)

    def t(fn):
        fn()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(3):
            fn()
        e.record()
        e.synchronize()
        return s.elapsed_time(e)

    t_own = t(lambda: gemm_sm100.matmul_nt(a, b, out=out))
    t_lib = t(lambda: torch.matmul(a, b.t(), out=out))
    choice = "sm100" if t_own <= t_lib else "cublas"
    _tuned[key] = choice
    return choice


def tuning_table():
    return dict(_tuned)


def matmul_nt(a, b):
    """a [M, K] @ b[N, K]^T."""
    if _sm100_usable(a, b, True) and _pick(a, b) == "sm100":
        from deepspeed_b200.ops.kernels import gemm_sm100
        _report(a.shape[0] * a.shape[1] * b.shape[0])
        return gemm_sm100.matmul_nt(a, b)
    return torch.matmul(a, b.t())


def matmul_nn(a, b):
    """a [M, K] @ b[K, N]."""
    if _sm100_usable(a, b, False):
        from deepspeed_b200.ops.kernels import gemm_sm100
        _report(a.shape[0] * a.shape[1] * b.shape[1])
        return gemm_sm100.matmul_nn(a, b)
    return torch.matmul(a, b)


def _report(macs):
    """Make the ctypes-launched GEMM visible to an active FlopsProfiler (ATen calls are counted by dispatch)."""
    from deepspeed_b200.profiling.flops_profiler import profiler as _p
    if _p._ACTIVE:
        _p.add_flops(2 * macs, macs)
