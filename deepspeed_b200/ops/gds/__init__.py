"""GPUDirect-Storage handle: NVMe <-> HBM without a host bounce when cuFile is usable, otherwise a pinned
bounce-buffer pipeline on the aio engine with the same API.

API parity: reference ``gds_handle`` (``csrc/gds/py_lib/deepspeed_py_gds_handle.cpp`` N11): the aio_handle
surface + ``new_pinned_device_tensor`` / ``free_pinned_device_tensor`` / ``pin_device_tensor`` /
``unpin_device_tensor`` (cuFileBufRegister).  cuFile is reached through ``ctypes`` on ``libcufile.so`` so there
is no build-time dependency.
"""
import ctypes
import os

import torch

from deepspeed_b200.ops.aio import aio_handle, AIO_DEFAULT_BLOCK_SIZE, AIO_DEFAULT_QUEUE_DEPTH


class _CUfileDescr(ctypes.Structure):
    _fields_ = [("type", ctypes.c_int), ("fd", ctypes.c_int), ("pad", ctypes.c_byte * 24), ("fs_ops", ctypes.c_void_p)]


class _CUfileError(ctypes.Structure):
    _fields_ = [("err", ctypes.c_int), ("cu_err", ctypes.c_int)]


_cufile = None
_cufile_state = None


def _load_cufile():
    """-> lib or None.  Only used on a CUDA device with a working nvidia-fs / compat mode."""
    global _cufile, _cufile_state
    if _cufile_state is not None:
        return _cufile
    _cufile_state = False
    if not torch.cuda.is_available() or os.environ.get("DSB200_GDS", "1") == "0":
        return None
    try:
        lib = ctypes.CDLL("libcufile.so.0")
        lib.cuFileDriverOpen.restype = _CUfileError
        lib.cuFileHandleRegister.restype = _CUfileError
        lib.cuFileHandleRegister.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(_CUfileDescr)]
        lib.cuFileHandleDeregister.argtypes = [ctypes.c_void_p]
        lib.cuFileBufRegister.restype = _CUfileError
        lib.cuFileBufRegister.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]
        lib.cuFileBufDeregister.restype = _CUfileError
        lib.cuFileBufDeregister.argtypes = [ctypes.c_void_p]
        for f in (lib.cuFileRead, lib.cuFileWrite):
            f.restype = ctypes.c_ssize_t
            f.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int64, ctypes.c_int64]
        st = lib.cuFileDriverOpen()
        if st.err != 0:
            return None
        _cufile, _cufile_state = lib, True
    except OSError:
        _cufile = None
    return _cufile


class gds_handle(aio_handle):

    def __init__(self, block_size=AIO_DEFAULT_BLOCK_SIZE, queue_depth=AIO_DEFAULT_QUEUE_DEPTH, single_submit=False,
                 overlap_events=True, intra_op_parallelism=1):
        super().__init__(block_size, queue_depth, single_submit, overlap_events, intra_op_parallelism)
        self._cf = _load_cufile()
        self._registered = {}
        self._bounce = None
        self._pending_dev = []

    @property
    def direct(self) -> bool:
        return self._cf is not None

    # ---- device buffers
    def new_pinned_device_tensor(self, num_elem, example_tensor):
        t = torch.empty(int(num_elem), dtype=example_tensor.dtype, device="cuda")
        self.pin_device_tensor(t)
        return t

    def free_pinned_device_tensor(self, tensor):
        self.unpin_device_tensor(tensor)
        return True

    def pin_device_tensor(self, tensor):
        if self._cf is not None and tensor.data_ptr() not in self._registered:
            st = self._cf.cuFileBufRegister(ctypes.c_void_p(tensor.data_ptr()), tensor.numel() * tensor.element_size(), 0)
            self._registered[tensor.data_ptr()] = st.err == 0
        return True

    def unpin_device_tensor(self, tensor):
        if self._cf is not None and self._registered.pop(tensor.data_ptr(), False):
            self._cf.cuFileBufDeregister(ctypes.c_void_p(tensor.data_ptr()))
        return True

    # ---- I/O (device tensors take the direct / bounce path, host tensors the plain aio path)
    def _direct(self, buffer, filename, write, file_offset):
        flags = (os.O_WRONLY | os.O_CREAT) if write else os.O_RDONLY
        fd = os.open(filename, flags | getattr(os, "O_DIRECT", 0), 0o644)
        try:
            d = _CUfileDescr(type=1, fd=fd)
            h = ctypes.c_void_p()
            st = self._cf.cuFileHandleRegister(ctypes.byref(h), ctypes.byref(d))
            if st.err != 0:
                return None
            n = buffer.numel() * buffer.element_size()
            fn = self._cf.cuFileWrite if write else self._cf.cuFileRead
            rc = fn(h, ctypes.c_void_p(buffer.data_ptr()), n, int(file_offset), 0)
            self._cf.cuFileHandleDeregister(h)
            return rc if rc >= 0 else None
        finally:
            os.close(fd)

    def _bounce_buf(self, nbytes):
        if self._bounce is None or self._bounce.numel() < nbytes:
            self._bounce = torch.empty(nbytes, dtype=torch.uint8, pin_memory=torch.cuda.is_available())
        return self._bounce[:nbytes]

    def pread(self, buffer, filename, validate=False, async_op=False, file_offset=0):
        if buffer.device.type != "cuda":
            return super().pread(buffer, filename, validate, async_op, file_offset)
        if self._cf is not None:
            torch.cuda.current_stream().synchronize()
            rc = self._direct(buffer, filename, False, file_offset)
            if rc is not None:
                return 0 if async_op else 1  # the reference convention: completed requests of a blocking call
        n = buffer.numel() * buffer.element_size()
        host = self._bounce_buf(n)
        super().pread(host, filename, validate, False, file_offset)
        buffer.view(torch.uint8).reshape(-1).copy_(host, non_blocking=True)
        if not async_op:
            torch.cuda.current_stream().synchronize()
        return 0 if async_op else 1

    def pwrite(self, buffer, filename, validate=False, async_op=False, file_offset=0):
        if buffer.device.type != "cuda":
            return super().pwrite(buffer, filename, validate, async_op, file_offset)
        if self._cf is not None:
            torch.cuda.current_stream().synchronize()
            rc = self._direct(buffer, filename, True, file_offset)
            if rc is not None:
                return 0 if async_op else 1
        n = buffer.numel() * buffer.element_size()
        host = self._bounce_buf(n)
        host.copy_(buffer.view(torch.uint8).reshape(-1))
        torch.cuda.current_stream().synchronize()
        super().pwrite(host, filename, validate, False, file_offset)
        return 0 if async_op else 1


class GDSBuilder:

    def load(self, verbose=False):
        import sys
        return sys.modules[__name__]

    def is_compatible(self, verbose=False):
        return True
