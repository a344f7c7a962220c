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
    if torch.cuda.is_current_stream_capturing():
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

    cands = {"sm100": t(lambda: gemm_sm100.matmul_nt(a, b, out=out)),
             "cublas": t(lambda: torch.matmul(a, b.t(), out=out))}
    if _two_cta_ok():
        cands["sm100_2cta"] = t(lambda: gemm_sm100.matmul_nt_2cta(a, b, out=out))
    if _backend == "sm100":
        cands.pop("cublas")
    choice = min(cands, key=cands.get)
    _tuned[key] = choice
    return choice


_2cta_ok = None


def _two_cta_ok():
    """One-time numerical self check of the CTA-pair kernel (disabled with DSB200_GEMM_2CTA=0)."""
    global _2cta_ok
    if _2cta_ok is None:
        if os.environ.get("DSB200_GEMM_2CTA", "1") == "0":
            _2cta_ok = False
            return False
        try:
            from deepspeed_b200.ops.kernels import gemm_sm100
            g = torch.Generator(device="cuda").manual_seed(0)
            a = torch.randn(640, 320, device="cuda", generator=g).bfloat16()
            b = torch.randn(768, 320, device="cuda", generator=g).bfloat16()
            c = gemm_sm100.matmul_nt_2cta(a, b)
            ref = a.float() @ b.float().t()
            torch.cuda.synchronize()
            _2cta_ok = bool((c.float() - ref).abs().max() < 0.05 * ref.abs().max() + 0.5) and bool(torch.isfinite(c).all())
        except Exception:
            _2cta_ok = False
    return _2cta_ok


def tuning_table():
    return dict(_tuned)


def matmul_nt(a, b):
    """a [M, K] @ b[N, K]^T."""
    if _sm100_usable(a, b, True):
        choice = _pick(a, b)
        if choice != "cublas":
            from deepspeed_b200.ops.kernels import gemm_sm100
            _report(a.shape[0] * a.shape[1] * b.shape[0])
            return gemm_sm100.matmul_nt_2cta(a, b) if choice == "sm100_2cta" else gemm_sm100.matmul_nt(a, b)
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
