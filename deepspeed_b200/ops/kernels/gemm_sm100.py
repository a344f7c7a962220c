"""Python face of ``csrc/cuda/gemm_sm100.cu`` -- the hand-written tcgen05 / TMEM / TMA GEMM."""
import ctypes

import torch

from deepspeed_b200.ops import native as N

_checked = None


def supports(a, b, nt=True) -> bool:
    """bf16, 2-D, K-major operands (``a`` [M,K], ``b`` [N,K] when ``nt``), K % 8 == 0, 16-byte aligned rows."""
    if not nt:
        return False  # NN / TN use the library path until the MN-major descriptors land
    if a.dtype != torch.bfloat16 or b.dtype != torch.bfloat16 or a.dim() != 2 or b.dim() != 2:
        return False
    if a.stride(1) != 1 or b.stride(1) != 1 or a.shape[1] != b.shape[1]:
        return False
    K = a.shape[1]
    if K % 8 or a.stride(0) % 8 or b.stride(0) % 8 or a.data_ptr() % 16 or b.data_ptr() % 16:
        return False
    return a.shape[0] >= 128 and b.shape[0] >= 256 and K >= 64


def matmul_nt(a, b, out=None, sms=0):
    """``a [M,K] @ b[N,K]^T -> [M,N]`` bf16."""
    M, K = a.shape
    Nn = b.shape[0]
    if out is None:
        out = torch.empty(M, Nn, dtype=torch.bfloat16, device=a.device)
    rc = N.cuda().dsb_gemm_nt_bf16(N.ptr(a), N.ptr(b), N.ptr(out), M, Nn, K, a.stride(0), b.stride(0), out.stride(0), sms,
                                   N.stream())
    N.check(rc, "gemm_nt_bf16")
    return out


def supports_grouped(x, w, offsets) -> bool:
    """bf16 rows [R, K] (row stride % 8), stacked weights [E, N, K] contiguous, int32 device offsets [E + 1]."""
    return (x.is_cuda and x.dtype == torch.bfloat16 and torch.is_tensor(w) and w.dtype == torch.bfloat16 and w.dim() == 3
            and w.is_contiguous() and x.dim() == 2 and x.stride(1) == 1 and x.shape[1] == w.shape[2] and x.shape[1] % 8 == 0
            and w.shape[1] % 8 == 0 and x.stride(0) % 8 == 0 and w.shape[0] <= 256 and torch.is_tensor(offsets)
            and offsets.is_cuda and offsets.dtype == torch.int32 and x.data_ptr() % 16 == 0 and w.data_ptr() % 16 == 0
            and x.shape[1] >= 64)


def grouped_matmul_nt(x, w, offsets, out=None, sms=0):
    """Grouped (MoE) GEMM: ``out[r] = x[r] @ w[e]^T`` for ``offsets[e] <= r < offsets[e+1]`` in ONE persistent tcgen05
    launch; ``offsets`` stays on the device (no host sync, CUDA-graph capturable)."""
    R, K = x.shape
    E, Nn, _ = w.shape
    if out is None:
        out = torch.empty(R, Nn, dtype=torch.bfloat16, device=x.device)
    rc = N.cuda().dsb_gemm_grouped_nt_bf16(N.ptr(x), N.ptr(w), N.ptr(out), N.ptr(offsets), R, E, Nn, K, x.stride(0), out.stride(0),
                                           sms, N.stream())
    N.check(rc, "gemm_grouped_nt_bf16")
    return out


def matmul_nt_2cta(a, b, out=None, sms=0):
    """Same contract as :func:`matmul_nt`, CTA-pair kernel (``tcgen05.mma.cta_group::2``, 256x256 tiles)."""
    M, K = a.shape
    Nn = b.shape[0]
    if out is None:
        out = torch.empty(M, Nn, dtype=torch.bfloat16, device=a.device)
    rc = N.cuda().dsb_gemm_nt_bf16_2cta(N.ptr(a), N.ptr(b), N.ptr(out), M, Nn, K, a.stride(0), b.stride(0), out.stride(0),
                                        sms, N.stream())
    N.check(rc, "gemm_nt_bf16_2cta")
    return out


EPI_STORE, EPI_ACCUM, EPI_SWIGLU, EPI_DSWIGLU = 0, 1, 2, 3


def matmul_2cta(a, b, a_mn: bool, b_mn: bool, out=None, sms=0, epi=EPI_STORE, aux=None, out2=None, inter=0, group_m=8):
    """CTA-pair kernel with selectable operand majorness and fused epilogue.

    ``a_mn=False``: ``a`` is [M, K];  ``a_mn=True``: ``a`` is [K, M] (the kernel multiplies by its transpose).
    ``b_mn=False``: ``b`` is [N, K];  ``b_mn=True``: ``b`` is [K, N].
    Result ``[M, N]`` bf16.  Rows may be strided (views of larger buffers).
    ``epi``: ``EPI_ACCUM`` adds into ``out``; ``EPI_SWIGLU`` -- ``b`` is ``[2*inter, K]`` (gate rows, then up rows), ``out``
    is ``[M, inter]`` = ``silu(gate) * up`` and ``out2`` (optional ``[M, 2*inter]``) receives gate|up; ``EPI_DSWIGLU`` --
    ``aux`` is the saved gate|up ``[M, 2*inter]``, ``out`` the dgate half and ``out2`` the whole ``[M, 2*inter]`` dgate|dup.
    ``group_m``: m-blocks per rasterisation super-group (L2 locality)."""
    M, K = (a.shape[1], a.shape[0]) if a_mn else (a.shape[0], a.shape[1])
    if epi == EPI_SWIGLU:
        Nn = inter
    else:
        Nn = b.shape[1] if b_mn else b.shape[0]
    if out is None:
        assert epi != EPI_ACCUM
        out = torch.empty(M, Nn, dtype=torch.bfloat16, device=a.device)
    rc = N.cuda().dsb_gemm_bf16_2cta_ex(N.ptr(a), N.ptr(b), N.ptr(out), M, Nn, K, a.stride(0), b.stride(0), out.stride(0),
                                        int(a_mn), int(b_mn), int(epi), N.ptr(aux), aux.stride(0) if aux is not None else 0,
                                        N.ptr(out2), out2.stride(0) if out2 is not None else 0, int(inter), int(group_m),
                                        sms, N.stream())
    N.check(rc, "gemm_bf16_2cta_ex")
    return out


def supports_2cta(a, b, a_mn, b_mn, out=None) -> bool:
    ok = lambda t: (t.is_cuda and t.dtype == torch.bfloat16 and t.dim() == 2 and t.stride(1) == 1 and t.stride(0) % 8 == 0
                    and t.data_ptr() % 16 == 0 and t.shape[1] % 8 == 0)
    return ok(a) and ok(b) and (out is None or ok(out))


def matmul_nn(a, b, out=None):
    """a [M, K] @ b [K, N]."""
    return matmul_2cta(a, b, False, True, out=out)


def matmul_tn(a, b, out=None):
    """a[K, M]^T @ b[K, N]."""
    return matmul_2cta(a, b, True, True, out=out)


def matmul_nt_allgather(a, b_view, local_full, peers, flags, shard_bytes, chunk_bytes, b_offset_bytes, world, rank, epoch,
                        comm_ctas=16, out=None, sms=0):
    """Fused all-gather + GEMM (see kernel header).  ``b_view`` is the [N,K] view of the weight inside
    ``local_full`` (the unit's gathered buffer, being filled by this very kernel)."""
    M, K = a.shape
    Nn = b_view.shape[0]
    if out is None:
        out = torch.empty(M, Nn, dtype=torch.bfloat16, device=a.device)
    arr = (ctypes.c_void_p * world)(*[ctypes.c_void_p(p) for p in peers])
    rc = N.cuda().dsb_gemm_nt_bf16_allgather(N.ptr(a), N.ptr(b_view), N.ptr(out), M, Nn, K, a.stride(0), b_view.stride(0),
                                             out.stride(0), arr, N.ptr(local_full), N.ptr(flags),
                                             ctypes.c_int64(shard_bytes), ctypes.c_int64(chunk_bytes),
                                             ctypes.c_int64(b_offset_bytes), world, rank, comm_ctas,
                                             ctypes.c_uint32(epoch), sms, N.stream())
    N.check(rc, "gemm_nt_bf16_allgather")
    return out


def self_check() -> bool:
    """One-time numerical self test against torch.matmul on this device."""
    global _checked
    if _checked is not None:
        return _checked
    try:
        torch.manual_seed(0)
        a = torch.randn(384, 320, device="cuda", dtype=torch.bfloat16)
        b = torch.randn(768, 320, device="cuda", dtype=torch.bfloat16)
        c = matmul_nt(a, b)
        ref = (a.float() @ b.float().t())
        torch.cuda.synchronize()
        err = (c.float() - ref).abs().max().item()
        _checked = bool(err < 0.05 * ref.abs().max().item() + 0.5) and bool(torch.isfinite(c).all())
    except Exception:
        _checked = False
    return _checked
