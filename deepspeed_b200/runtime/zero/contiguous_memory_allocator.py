"""Defragmenting sub-allocator over one flat buffer (reference ``runtime/zero/contiguous_memory_allocator.py``).

Used by activation checkpointing (``contiguous_memory_optimization``) and by callers that want parameter storage carved
from one arena.  Bookkeeping is an address-sorted list of live blocks; free space is the gaps between them.  When no gap
fits a request but the total free space does, live blocks are slid to the front (``Tensor.set_`` / ``param.data``
rebinding keeps every handed-out tensor valid) and the request is served from the tail.
"""
import bisect

import torch

from deepspeed_b200.utils import logger


class _Block:
    __slots__ = ("addr", "size", "tensor", "params")

    def __init__(self, addr, size, tensor):
        self.addr, self.size, self.tensor, self.params = addr, size, tensor, []


class ContiguousMemoryAllocator:

    def __init__(self, size, dtype, device):
        self.buffer = torch.zeros(size, dtype=dtype, device=device)
        self.total_size = size
        self.total_free = size
        self.max_allocated = 0
        self.count = 0
        self._blocks = []  # sorted by addr
        self._by_id = {}  # id(tensor) -> _Block

    # -- queries ----------------------------------------------------------------------------------------------------
    def _gaps(self):
        """Yield (addr, size) of every free interval in address order."""
        cur = 0
        for b in self._blocks:
            if b.addr > cur:
                yield cur, b.addr - cur
            cur = b.addr + b.size
        if cur < self.total_size:
            yield cur, self.total_size - cur

    @property
    def largest_contiguous(self):
        return max((s for _, s in self._gaps()), default=0)

    @property
    def tensor_map(self):
        return {i: b.tensor for i, b in self._by_id.items()}

    def print_allocation(self, resolution=200):
        cell = max(1, self.total_size // resolution)
        line = ["."] * ((self.total_size + cell - 1) // cell)
        for b in self._blocks:
            for i in range(b.addr // cell, min(len(line), (b.addr + b.size - 1) // cell + 1)):
                line[i] = "x"
        logger.info("".join(line))

    # -- allocate / release -------------------------------------------------------------------------------------------
    def allocate_tensor(self, size):
        assert size <= self.total_free, "Not enough memory in buffer. Allocation failed"
        addr = next((a for a, s in self._gaps() if s >= size), None)
        if addr is None:
            self._defragment_memory()
            addr = next(a for a, s in self._gaps() if s >= size)
        t = self.buffer.narrow(0, addr, size)
        blk = _Block(addr, size, t)
        bisect.insort(self._blocks, blk, key=lambda b: b.addr)
        self._by_id[id(t)] = blk
        self.total_free -= size
        self.max_allocated = max(self.max_allocated, self.total_size - self.total_free)
        self.count += 1
        return t

    def assign_to_param(self, tensor, param, numel, shape):
        blk = self._by_id.get(id(tensor))
        assert blk is not None, "No such tensor allocated by the allocator."
        assert tensor.numel() >= numel, "Tensor buffer is not large enough"
        assert not blk.params, "This tensor has already been assigned to a param"
        blk.params.append((param, numel, tuple(shape)))
        param.data = tensor.narrow(0, 0, numel).view(shape)
        param.contiguous_tensor_id = id(tensor)

    def release_tensor(self, tensor):
        self.release_tensor_with_id(id(tensor))

    def release_tensor_with_id(self, tensor_id):
        blk = self._by_id.pop(tensor_id, None)
        assert blk is not None, "Invalid tensor id"
        self._blocks.remove(blk)
        for param, _, _ in blk.params:
            param.data = torch.empty(0, dtype=param.dtype, device=param.device)
        self.total_free += blk.size

    # -- compaction ---------------------------------------------------------------------------------------------------
    def _defragment_memory(self):
        """Slide every live block to the lowest free address (ascending order, so sources are never overwritten before
        they are read; overlapping moves go through a clone)."""
        cur = 0
        for b in self._blocks:
            if b.addr != cur:
                src = self.buffer.narrow(0, b.addr, b.size)
                dst = self.buffer.narrow(0, cur, b.size)
                dst.copy_(src.clone() if cur + b.size > b.addr else src)
                b.addr = cur
                b.tensor.set_(dst.untyped_storage(), dst.storage_offset(), dst.shape, dst.stride())  # same id, new home
                for param, numel, shape in b.params:
                    param.data = b.tensor.narrow(0, 0, numel).view(shape)
            cur += b.size


def print_rank_0(message, debug=False, force=False):
    from deepspeed_b200 import comm as dist
    if (debug or force) and (not dist.is_initialized() or dist.get_rank() == 0):
        print(message)
