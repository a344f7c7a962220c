"""Host-side ragged helpers backed by ``csrc/cpu/host_utils.cpp`` (reference ``inference/v2/ragged/csrc`` N9c):
attention-atom construction and a native block free-list."""
import ctypes

import torch

from deepspeed_b200.ops import native as N


def build_atoms(seq_new_tokens, seq_seen_tokens, seq_token_start, seq_block_table_off, q_block, kv_block, max_atoms=4096):
    """One atom per (sequence, q-block): int32 rows [seq, q_token_start, q_len, kv_blocks_visible, kv_len, block_table_off]."""
    lib = N.cpu()
    lib.dsb_build_atoms.restype = ctypes.c_int64
    n = seq_new_tokens.numel()
    out = torch.empty(max_atoms, 6, dtype=torch.int32)
    args = [t.to(torch.int32).contiguous() for t in (seq_new_tokens, seq_seen_tokens, seq_token_start, seq_block_table_off)]
    cnt = lib.dsb_build_atoms(*(ctypes.c_void_p(t.data_ptr()) for t in args), n, q_block, kv_block,
                              ctypes.c_void_p(out.data_ptr()), ctypes.c_int64(max_atoms))
    if cnt < 0:
        raise RuntimeError("atom buffer too small")
    return out[:cnt]


class NativeBlockAllocator:

    def __init__(self, n_blocks):
        lib = N.cpu()
        lib.dsb_blockalloc_create.restype = ctypes.c_void_p
        lib.dsb_blockalloc_create.argtypes = [ctypes.c_int32]
        lib.dsb_blockalloc_destroy.argtypes = [ctypes.c_void_p]
        lib.dsb_blockalloc_free_count.argtypes = [ctypes.c_void_p]
        lib.dsb_blockalloc_allocate.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p]
        lib.dsb_blockalloc_free.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32]
        self._lib, self._h, self.total_blocks = lib, lib.dsb_blockalloc_create(n_blocks), n_blocks

    @property
    def free_blocks(self):
        return self._lib.dsb_blockalloc_free_count(self._h)

    def allocate(self, n):
        out = torch.empty(n, dtype=torch.int32)
        if self._lib.dsb_blockalloc_allocate(self._h, n, ctypes.c_void_p(out.data_ptr())) < 0:
            raise ValueError(f"Not enough free blocks in the KV-cache to allocate {n} blocks")
        return out

    def free(self, blocks):
        b = torch.as_tensor(blocks, dtype=torch.int32).reshape(-1).contiguous()
        if self._lib.dsb_blockalloc_free(self._h, ctypes.c_void_p(b.data_ptr()), b.numel()) < 0:
            raise ValueError("invalid or already-free block")

    def __del__(self):
        try:
            self._lib.dsb_blockalloc_destroy(self._h)
        except Exception:
            pass
