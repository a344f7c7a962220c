"""Async file I/O handle over the native engine (``csrc/cpu/aio.cpp``: Linux AIO via raw syscalls, O_DIRECT,
worker threads that split a request into aligned slices).

API parity: reference ``aio_handle`` (``csrc/aio/py_lib/deepspeed_py_io_handle.cpp`` N10): ``sync_pread/pwrite``,
``async_pread/pwrite``, ``pread/pwrite(buffer, path, validate, async, file_offset)``, ``wait``,
``new_cpu_locked_tensor``/``free_cpu_locked_tensor``, ``get_block_size`` & friends.
"""
import ctypes
import os

import torch

from deepspeed_b200.ops import native as N

AIO_DEFAULT_BLOCK_SIZE = 1 << 20
AIO_DEFAULT_QUEUE_DEPTH = 32
AIO_DEFAULT_INTRA_OP_PARALLELISM = 1


def _lib():
    lib = N.cpu()
    if not getattr(lib, "_aio_typed", False):
        lib.dsb_aio_create.restype = ctypes.c_void_p
        lib.dsb_aio_create.argtypes = [ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        lib.dsb_aio_destroy.argtypes = [ctypes.c_void_p]
        for f in (lib.dsb_aio_pread, lib.dsb_aio_pwrite):
            f.restype = ctypes.c_int64
            f.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_char_p, ctypes.c_int64, ctypes.c_int]
        lib.dsb_aio_wait.restype = ctypes.c_int64
        lib.dsb_aio_wait.argtypes = [ctypes.c_void_p]
        lib.dsb_aio_alloc_locked.restype = ctypes.c_void_p
        lib.dsb_aio_alloc_locked.argtypes = [ctypes.c_int64]
        lib.dsb_aio_free_locked.argtypes = [ctypes.c_void_p, ctypes.c_int64]
        lib.dsb_file_size.restype = ctypes.c_int64
        lib.dsb_file_size.argtypes = [ctypes.c_char_p]
        lib.dsb_parallel_memcpy.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int]
        lib._aio_typed = True
    return lib


class aio_handle:

    def __init__(self, block_size=AIO_DEFAULT_BLOCK_SIZE, queue_depth=AIO_DEFAULT_QUEUE_DEPTH, single_submit=False,
                 overlap_events=True, intra_op_parallelism=AIO_DEFAULT_INTRA_OP_PARALLELISM):
        self._lib = _lib()
        self._block_size, self._queue_depth = int(block_size), int(queue_depth)
        self._single_submit, self._overlap_events = bool(single_submit), bool(overlap_events)
        self._threads = int(intra_op_parallelism)
        self._h = self._lib.dsb_aio_create(self._block_size, self._queue_depth, int(self._single_submit),
                                           int(self._overlap_events), self._threads)
        if not self._h:
            raise RuntimeError("failed to create the async I/O context")
        self._inflight = []
        self._dev_reads = []  # (pinned host, device) pairs of asynchronous reads into device tensors
        # keep tensors alive until wait()
        self._locked = {}

    def __del__(self):
        try:
            if self._h:
                self._lib.dsb_aio_wait(self._h)
                self._lib.dsb_aio_destroy(self._h)
                self._h = None
        except Exception:
            pass

    # ---- config getters
    def get_block_size(self):
        return self._block_size

    def get_queue_depth(self):
        return self._queue_depth

    def get_single_submit(self):
        return self._single_submit

    def get_overlap_events(self):
        return self._overlap_events

    def get_intra_op_parallelism(self):
        return self._threads

    get_thread_count = get_intra_op_parallelism

    def get_alignment(self):
        return 4096

    # ---- I/O
    def _io(self, fn, buffer, filename, is_async, file_offset):
        """Reference return convention (``py_lib/deepspeed_py_io_handle.cpp``): a blocking call returns the number of
        completed requests (1), an asynchronous one returns 0 and ``wait()`` reports the count."""
        assert buffer.is_contiguous(), "aio buffers must be contiguous"
        is_read = fn is self._lib.dsb_aio_pread
        dev_buffer = None
        if buffer.device.type != "cpu":
            # device tensors go through a pinned bounce buffer (the GDS handle in ``ops/gds`` is the direct path)
            dev_buffer = buffer
            buffer = torch.empty(dev_buffer.shape, dtype=dev_buffer.dtype, device="cpu", pin_memory=torch.cuda.is_available())
            if not is_read:
                buffer.copy_(dev_buffer)
        n = buffer.numel() * buffer.element_size()
        rc = fn(self._h, ctypes.c_void_p(buffer.data_ptr()), n, os.fsencode(filename), int(file_offset), int(is_async))
        if rc < 0:
            raise OSError(-rc, f"aio request on {filename} failed: {os.strerror(-rc)}")
        if is_async:
            self._inflight.append(buffer)
            if dev_buffer is not None and is_read:
                self._dev_reads.append((buffer, dev_buffer))
            return 0
        if dev_buffer is not None and is_read:
            dev_buffer.copy_(buffer)
        return 1

    def pread(self, buffer, filename, validate=False, async_op=False, file_offset=0):
        return self._io(self._lib.dsb_aio_pread, buffer, filename, async_op, file_offset)

    def pwrite(self, buffer, filename, validate=False, async_op=False, file_offset=0):
        return self._io(self._lib.dsb_aio_pwrite, buffer, filename, async_op, file_offset)

    def sync_pread(self, buffer, filename, file_offset=0):
        return self.pread(buffer, filename, False, False, file_offset)

    def sync_pwrite(self, buffer, filename, file_offset=0):
        return self.pwrite(buffer, filename, False, False, file_offset)

    def async_pread(self, buffer, filename, file_offset=0):
        return self.pread(buffer, filename, False, True, file_offset)

    def async_pwrite(self, buffer, filename, file_offset=0):
        return self.pwrite(buffer, filename, False, True, file_offset)

    def read(self, buffer, filename, validate=False):
        return self.sync_pread(buffer, filename)

    def write(self, buffer, filename, validate=False):
        return self.sync_pwrite(buffer, filename)

    def wait(self):
        rc = self._lib.dsb_aio_wait(self._h)
        n = len(self._inflight)
        self._inflight.clear()
        if rc < 0:
            raise OSError(-rc, f"aio wait failed: {os.strerror(-rc)}")
        for host, dev in self._dev_reads:
            dev.copy_(host)
        self._dev_reads.clear()
        return n

    # ---- locked host tensors
    def new_cpu_locked_tensor(self, num_elem, example_tensor):
        dtype = example_tensor.dtype
        nbytes = int(num_elem) * dtype.itemsize
        ptr = self._lib.dsb_aio_alloc_locked(nbytes)
        if not ptr:
            raise MemoryError(f"cannot allocate {nbytes} locked bytes")
        buf = (ctypes.c_uint8 * nbytes).from_address(ptr)
        t = torch.frombuffer(buf, dtype=dtype, count=int(num_elem))
        self._locked[t.data_ptr()] = (ptr, nbytes, buf)
        return t

    def free_cpu_locked_tensor(self, tensor):
        ent = self._locked.pop(tensor.data_ptr(), None)
        if ent is None:
            return False
        self._lib.dsb_aio_free_locked(ent[0], ent[1])
        return True


def file_size(path):
    return _lib().dsb_file_size(os.fsencode(path))


def parallel_memcpy(dst, src, threads=8):
    n = dst.numel() * dst.element_size()
    assert n == src.numel() * src.element_size()
    _lib().dsb_parallel_memcpy(ctypes.c_void_p(dst.data_ptr()), ctypes.c_void_p(src.data_ptr()), n, threads)


class AsyncIOBuilder:
    """Compat shim: ``AsyncIOBuilder().load().aio_handle(...)``."""

    def load(self, verbose=False):
        import sys
        return sys.modules[__name__]

    def is_compatible(self, verbose=False):
        try:
            _lib()
            return True
        except Exception:
            return False
