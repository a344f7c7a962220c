"""Read-ahead / write-behind NVMe optimizer-state swapper (reference
``runtime/swap_tensor/pipelined_optimizer_swapper.py``): while the step works on window *i*, window *i+1* is being read
and window *i-1* written (``FlatStateSwapper`` with its pipeline switched on, one extra staging window)."""
from .optimizer_utils import FlatStateSwapper


class PipelinedOptimizerSwapper(FlatStateSwapper):

    def __init__(self, swap_config, aio_config, base_folder, optimizer=None, largest_numel=None, device=None, dtype=None,
                 timers=None, rank=0):
        import torch
        super().__init__(swap_config, aio_config, base_folder, rank, dtype or torch.float32)
        self.pipeline = True
        self.optimizer, self.timers = optimizer, timers

    def swap_in_optimizer_state(self, flat_opt, start, end, next_range=None):
        """Resident view of [start, end); kicks off the read of ``next_range`` so it overlaps this window's math."""
        if next_range is not None:
            self.prefetch(flat_opt, *next_range)
        return {n: t[start:end] for n, t in flat_opt.state.items()}

    def swap_out_optimizer_state(self, flat_opt, async_swap=True):
        self.flush(flat_opt, wait=not async_swap)


class OptimizerSwapOp:
    """One queued unit of pipelined swap work (reference ``pipelined_optimizer_swapper.py:21``): the ranges being read /
    written and whether the corresponding aio batch has been waited for."""

    def __init__(self, aio_handle, read_op, param_info, allocated_buffers, state_buffers, num_ops):
        self.aio_handle, self.read_op, self.param_info = aio_handle, read_op, param_info
        self.allocated_buffers, self.state_buffers = allocated_buffers, state_buffers
        self.wait_required = True
        self.num_ops = num_ops

    def is_parameter(self, parameter):
        from .optimizer_utils import OptimizerSwapper
        return OptimizerSwapper.parameter_id(parameter) == self.param_info.param_id

    def wait(self):
        assert self.wait_required
        self.aio_handle.wait()
        self.wait_required = False


SYNC_SWAP_IN, ASYNC_SWAP_IN, SYNC_SWAP_OUT, ASYNC_SWAP_OUT = "sync_swap_in", "async_swap_in", "sync_swap_out", "async_swap_out"
