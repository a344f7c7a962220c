"""Decompose a ragged batch into attention *atoms*: (sequence, query-block) work items of at most ``q_block_size`` tokens with the KV range each one needs.  The fused paged-attention kernels here are token-parallel and do not consume atoms; the builder is provided for schedulers / external kernels that do.

Reference ``inference/v2/kernels/ragged_ops/atom_builder/atom_builder.py``."""
from typing import Tuple

import torch

from ...ds_kernel import DSKernelBase


class AtomBuilder(DSKernelBase):

    def __init__(self) -> None:
        pass

    def __call__(self, atoms: torch.Tensor, ragged_batch, q_block_size: int, kv_block_size: int) -> Tuple[torch.Tensor, int]:
        """``ragged_batch``: rows ``[start token, n tokens, seen tokens]`` per in-flight sequence (tensor / list) or an object
        with ``inflight_seq_descriptors``.  Fills ``atoms`` (int32 [max_atoms, 8]) with rows
        ``[seq_slot, q_start, q_len, kv_blocks, total_kv_len, global_q_pos, 0, 0]``; returns (atoms, n)."""
        seqs = ragged_batch.inflight_seq_descriptors(on_device=False) if hasattr(ragged_batch, "inflight_seq_descriptors") else ragged_batch
        n = 0
        host = atoms if not atoms.is_cuda else torch.zeros_like(atoms, device="cpu")
        for slot, row in enumerate(seqs.tolist() if torch.is_tensor(seqs) else seqs):
            start, n_tokens, seen = int(row[0]), int(row[1]), int(row[2])
            done = 0
            while done < n_tokens:
                q_len = min(q_block_size, n_tokens - done)
                total = seen + done + q_len
                host[n] = torch.tensor([slot, start + done, q_len, (total + kv_block_size - 1) // kv_block_size, total,
                                        seen + done, 0, 0], dtype=host.dtype)
                done += q_len
                n += 1
        if atoms.is_cuda:
            atoms.copy_(host, non_blocking=True)
        return atoms, n
