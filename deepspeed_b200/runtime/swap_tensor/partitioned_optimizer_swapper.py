"""Synchronous NVMe optimizer-state swapper (reference ``runtime/swap_tensor/partitioned_optimizer_swapper.py``).

Both reference swappers map onto ``optimizer_utils.FlatStateSwapper``: optimizer moments are flat NVMe-backed arrays
visited window by window during the step.  This variant never reads ahead: a window is read when the step reaches it
and written back before the next one is touched.
"""
from .optimizer_utils import FlatStateSwapper


class PartitionedOptimizerSwapper(FlatStateSwapper):

    def __init__(self, swap_config, aio_config, base_folder, optimizer=None, largest_numel=None, device=None, dtype=None,
                 timers=None, rank=0):
        import torch
        super().__init__(swap_config, aio_config, base_folder, rank, dtype or torch.float32)
        self.pipeline = False
        self.optimizer, self.timers = optimizer, timers

    def swap_in_optimizer_state(self, flat_opt, start, end):
        """Make [start, end) of every moment resident (blocking)."""
        return {n: t[start:end] for n, t in flat_opt.state.items()}

    def swap_out_optimizer_state(self, flat_opt, async_swap=False):
        self.flush(flat_opt, wait=not async_swap)
