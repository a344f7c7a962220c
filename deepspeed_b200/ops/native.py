"""ctypes front door to the native libraries.

``cuda()`` / ``cpu()`` return the loaded ``CDLL``; on a GPU box a missing CUDA library is a hard
error (ops never fall back silently).  Helper converters turn tensors into the (pointer, dtype-code,
stream) triples the C ABI expects.
"""
import ctypes
import functools

import torch

F32, F16, BF16, I8, U8, F8E4M3, F8E5M2 = 0, 1, 2, 3, 4, 5, 6
_DT = {
    torch.float32: F32,
    torch.float16: F16,
    torch.bfloat16: BF16,
    torch.int8: I8,
    torch.uint8: U8,
}
if hasattr(torch, "float8_e4m3fn"):
    _DT[torch.float8_e4m3fn] = F8E4M3
    _DT[torch.float8_e5m2] = F8E5M2


class NativeOpError(RuntimeError):
    pass


@functools.lru_cache(None)
def cuda() -> ctypes.CDLL:
    from deepspeed_b200.op_builder import CudaKernelsBuilder
    return CudaKernelsBuilder().load()


@functools.lru_cache(None)
def cpu() -> ctypes.CDLL:
    from deepspeed_b200.op_builder import CpuRuntimeBuilder
    return CpuRuntimeBuilder().load()


def cuda_available() -> bool:
    """True iff a CUDA device is present (then the native library MUST load)."""
    return torch.cuda.is_available()


def dt(t) -> int:
    d = t if isinstance(t, torch.dtype) else t.dtype
    try:
        return _DT[d]
    except KeyError:
        raise NativeOpError(f"dtype {d} is not supported by the native kernels")


def ptr(t):
    if t is None:
        return ctypes.c_void_p(0)
    return ctypes.c_void_p(t.data_ptr())


def stream(device=None):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


launch_count = 0  # native entry points invoked since import (each launches >= 1 kernel)


def check(rc: int, what: str):
    global launch_count
    launch_count += 1
    if rc == 0:
        return
    if rc == -1:
        raise NativeOpError(f"{what}: unsupported dtype combination")
    if rc == -2:
        raise NativeOpError(f"{what}: unsupported shape/alignment")
    raise NativeOpError(f"{what}: CUDA error {rc} ({_cuda_err(rc)})")


def _cuda_err(rc):
    try:
        rt = ctypes.CDLL("libcudart.so")
        rt.cudaGetErrorString.restype = ctypes.c_char_p
        return rt.cudaGetErrorString(rc).decode()
    except Exception:
        return "?"


c_f = ctypes.c_float
c_i = ctypes.c_int
c_i64 = ctypes.c_int64
