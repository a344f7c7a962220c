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


def _pick_kind(kind, a, b, out, lib_fn, own_fn):
    """Autotune NN / TN problems (backward GEMMs) between cuBLAS and the CTA-pair tcgen05 kernel."""
    key = (kind, ) + tuple(a.shape) + tuple(b.shape)
    choice = _tuned.get(key)
    if choice is not None:
        return choice
    if torch.cuda.is_current_stream_capturing():
        return "cublas"
    scratch = torch.empty_like(out)

    def t(fn):
        fn()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(3):
            fn()
        e.record()
        e.synchronize()
        return s.elapsed_time(e)

    t_lib = t(lambda: lib_fn(scratch))
    try:
        t_own = t(lambda: own_fn(scratch))
        ref = lib_fn(torch.empty_like(out)).float()
        good = bool((scratch.float() - ref).abs().max() <= 0.05 * ref.abs().max() + 0.5)
    except Exception:
        t_own, good = float("inf"), False
    choice = "sm100_2cta" if (good and t_own < t_lib and _backend != "cublas") else "cublas"
    _tuned[key] = choice
    return choice


def _own_ok(a, b, out):
    if _backend == "cublas" or not a.is_cuda or not _two_cta_ok():
        return False
    from deepspeed_b200.ops.kernels import gemm_sm100
    return gemm_sm100.supports_2cta(a, b, False, False, out)


def matmul_nn(a, b, out=None):
    """a [M, K] @ b[K, N]  (dX = dY @ W)."""
    if out is None:
        out = torch.empty(a.shape[0], b.shape[1], dtype=a.dtype, device=a.device)
    if _own_ok(a, b, out):
        from deepspeed_b200.ops.kernels import gemm_sm100
        if _pick_kind("nn", a, b, out, lambda o: torch.mm(a, b, out=o), lambda o: gemm_sm100.matmul_nn(a, b, out=o)) \
                == "sm100_2cta":
            _report(a.shape[0] * a.shape[1] * b.shape[1])
            return gemm_sm100.matmul_nn(a, b, out=out)
    return torch.mm(a, b, out=out)


def matmul_tn(a, b, out=None):
    """a[K, M]^T @ b[K, N]  (dW = dY^T @ X), optionally straight into ``out`` (a flat-gradient view)."""
    if out is None:
        out = torch.empty(a.shape[1], b.shape[1], dtype=a.dtype, device=a.device)
    if _own_ok(a, b, out):
        from deepspeed_b200.ops.kernels import gemm_sm100
        if _pick_kind("tn", a, b, out, lambda o: torch.mm(a.t(), b, out=o), lambda o: gemm_sm100.matmul_tn(a, b, out=o)) \
                == "sm100_2cta":
            _report(a.shape[0] * a.shape[1] * b.shape[1])
            return gemm_sm100.matmul_tn(a, b, out=out)
    return torch.mm(a.t(), b, out=out)


def _report(macs):
    """Make the ctypes-launched GEMM visible to an active FlopsProfiler (ATen calls are counted by dispatch)."""
    from deepspeed_b200.profiling.flops_profiler import profiler as _p
    if _p._ACTIVE:
        _p.add_flops(2 * macs, macs)
