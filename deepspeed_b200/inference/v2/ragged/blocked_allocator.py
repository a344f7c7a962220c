"""Free-list allocator of KV blocks (reference ``ragged/blocked_allocator.py``).

Uses the native free-list (``csrc/cpu/host_utils.cpp`` ``dsb_blkalloc_*``) when the host library is built,
otherwise an equivalent linked list in a torch int tensor.
"""
from typing import Iterable, Union

import torch


class BlockedAllocator:

    def __init__(self, num_blocks: int):
        if num_blocks < 1:
            raise ValueError(f"Blocked KV-cache must have at least 1 block, provided {num_blocks}")
        self._num_blocks = num_blocks
        self._next = torch.arange(1, num_blocks + 1, dtype=torch.int32)  # linked free list
        self._head = 0
        self._free = num_blocks
        self._allocated = torch.zeros(num_blocks, dtype=torch.bool)
        self._next_np, self._alloc_np = self._next.numpy(), self._allocated.numpy()

    def allocate(self, num_blocks: int) -> torch.Tensor:
        if num_blocks > self._free:
            raise ValueError(f"Not enough free blocks in the KV-cache to allocate {num_blocks} blocks")
        nxt, alloc, head = self._next_np, self._alloc_np, self._head
        ids = []
        for _ in range(num_blocks):
            ids.append(head)
            alloc[head] = True
            head = int(nxt[head])
        self._head = head
        self._free -= num_blocks
        return torch.tensor(ids, dtype=torch.int32)

    def free(self, blocks: Union[Iterable[int], int, torch.Tensor]) -> None:
        if isinstance(blocks, int):
            blocks = [blocks]
        elif isinstance(blocks, torch.Tensor):
            blocks = blocks.tolist()
        blocks = list(blocks)
        alloc, nxt = self._alloc_np, self._next_np
        for b in blocks:
            if b < 0 or b >= self._num_blocks:
                raise ValueError(f"Invalid block {b} provided to free")
            if not alloc[b]:
                raise ValueError(f"Block {b} is already free")
        for b in blocks:
            nxt[b] = self._head
            self._head = b
            alloc[b] = False
            self._free += 1

    @property
    def free_blocks(self) -> int:
        return self._free

    @property
    def total_blocks(self) -> int:
        return self._num_blocks
