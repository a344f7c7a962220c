"""Exact-size page-locked host tensors for the offload arenas.

``torch.empty(..., pin_memory=True)`` goes through torch's caching host allocator, which rounds every request up to the next
power of two: the four 4-byte-per-parameter arenas of the ZeRO-Offload tier (fp32 master, two Adam moments, reduced
gradient shard; reference ``stage3.py:854-856`` pins the same tensors) would take up to twice the memory they need --
enough to push a Llama-70B run over a node's host memory.  Large arenas therefore come from ``dsb_pinned_alloc``
(``csrc/cuda/symm_mem.cpp``): 2 MiB-aligned pages, first-touched by several threads, registered with the driver, exactly
the requested size.  ``tensor.is_pinned()`` is true for them and ``copy_(non_blocking=True)`` is asynchronous.
"""
import ctypes
import os
import weakref

import torch

THRESHOLD_BYTES = 256 << 20  # smaller buffers are fine in torch's allocator (and get recycled there)
_live = {}


def pinned_empty(numel: int, dtype: torch.dtype) -> torch.Tensor:
    """Uninitialised (zero-page backed) pinned 1-D tensor of exactly ``numel`` elements."""
    nbytes = int(numel) * torch.empty((), dtype=dtype).element_size()
    if not torch.cuda.is_available():
        return torch.empty(numel, dtype=dtype)
    if nbytes < THRESHOLD_BYTES:
        return torch.empty(numel, dtype=dtype, device="cpu", pin_memory=True)
    from deepspeed_b200.ops import native as N
    lib = N.cuda()
    lib.dsb_pinned_alloc.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int64, ctypes.c_int]
    lib.dsb_pinned_free.argtypes = [ctypes.c_void_p]
    ptr = ctypes.c_void_p()
    threads = max(1, min(16, (os.cpu_count() or 8) // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1")))))
    rc = lib.dsb_pinned_alloc(ctypes.byref(ptr), nbytes, threads)
    if rc != 0:
        raise MemoryError(f"pinned host allocation of {nbytes / 2**30:.1f} GiB failed (rc {rc})")
    buf = (ctypes.c_uint8 * nbytes).from_address(ptr.value)
    t = torch.frombuffer(buf, dtype=torch.uint8).view(dtype)
    addr = ptr.value
    _live[addr] = nbytes

    def _free(a=addr):
        if _live.pop(a, None) is not None:
            lib.dsb_pinned_free(ctypes.c_void_p(a))

    # the ctypes array is what torch.frombuffer keeps alive; when the last tensor over it dies, unregister + free the pages
    weakref.finalize(buf, _free)
    return t


def live_bytes() -> int:
    return sum(_live.values())
