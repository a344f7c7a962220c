"""Intra-host shared-memory collectives for CPU tensors (reference ``csrc/cpu/comm/shm.cpp`` N17): a fast path
for the host-side reductions of the offload tier when several ranks share a node."""
import ctypes
import os

import torch

from deepspeed_b200.ops import native as N

_DT = {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2}


class ShmComm:

    def __init__(self, rank: int, world: int, name: str = None, max_bytes: int = 1 << 24):
        lib = N.cpu()
        lib.dsb_shm_create.restype = ctypes.c_void_p
        lib.dsb_shm_create.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_int64]
        lib.dsb_shm_destroy.argtypes = [ctypes.c_void_p]
        lib.dsb_shm_barrier.argtypes = [ctypes.c_void_p]
        lib.dsb_shm_all_reduce.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int]
        lib.dsb_shm_all_gather.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64]
        self._lib = lib
        name = (name or f"shm_{os.environ.get('MASTER_PORT', '0')}").strip("/").replace("/", "_")
        self.rank, self.world, self.max_bytes = rank, world, max_bytes
        self._h = lib.dsb_shm_create(name.encode(), rank, world, max_bytes)
        if not self._h:
            raise RuntimeError("shared-memory communicator creation failed")

    def barrier(self):
        self._lib.dsb_shm_barrier(self._h)

    def all_reduce(self, t: torch.Tensor):
        assert t.device.type == "cpu" and t.is_contiguous() and t.dtype in _DT
        per = self.max_bytes // t.element_size()
        flat = t.view(-1)
        for s in range(0, flat.numel(), per):
            c = flat[s:s + per]
            rc = self._lib.dsb_shm_all_reduce(self._h, ctypes.c_void_p(c.data_ptr()), c.numel(), _DT[t.dtype])
            if rc != 0:
                raise RuntimeError(f"shm all_reduce failed rc={rc}")
        return t

    def all_gather(self, out: torch.Tensor, t: torch.Tensor):
        n = t.numel() * t.element_size()
        rc = self._lib.dsb_shm_all_gather(self._h, ctypes.c_void_p(t.data_ptr()), ctypes.c_void_p(out.data_ptr()), n)
        if rc != 0:
            raise RuntimeError(f"shm all_gather failed rc={rc}")
        return out

    def close(self):
        if self._h:
            self._lib.dsb_shm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
