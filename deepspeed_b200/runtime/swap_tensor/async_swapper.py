"""Fire-and-forget swap-out of a stream of tensors through a small pinned pool (reference
``runtime/swap_tensor/async_swapper.py:19 AsyncTensorSwapper``): used to spill gradients to NVMe."""
import torch

from .utils import AIO_ALIGNED_BYTES, SwapBufferPool


class AsyncTensorSwapper:

    def __init__(self, aio_handle, numel_alignment, timers=None):
        self.aio_handle = aio_handle
        self.numel_alignment = numel_alignment
        self.free_buffer_index, self.swapping_buffer_index, self.ready_buffer_index = [], [], []
        self.current_buffer_index = -1
        self.all_buffers = []
        self.max_numel = 0
        self.num_pending_swaps = 0
        self.num_elements_swapped = 0
        self.dtype = None

    def has_buffers(self):
        return len(self.all_buffers) > 0

    def add_buffers(self, buffer_list):
        assert not self.all_buffers and all(b.device.type == "cpu" for b in buffer_list)
        self.dtype = buffer_list[0].dtype
        from .utils import SwapBuffer
        self.all_buffers = [SwapBuffer(b) for b in buffer_list]
        self.free_buffer_index = list(range(len(buffer_list)))
        self.max_numel = max(b.numel() for b in buffer_list)

    def get_timer_names(self):
        return []

    def release_buffers(self):
        self._flush_buffers_until_complete()
        pinned = [b.buffer for b in self.all_buffers]
        self.all_buffers, self.free_buffer_index, self.current_buffer_index = [], [], -1
        self.num_elements_swapped = 0
        self.dtype = None
        return pinned

    def swap_out_tensors(self, tensor_list, path_list):
        for t, p in zip(tensor_list, path_list):
            self._swap_out_tensor(t, p)

    def _aligned(self, n):
        r = n % self.numel_alignment
        return n if r == 0 else n + self.numel_alignment - r

    def _swap_out_tensor(self, tensor, swap_path):
        assert self.all_buffers
        aligned = self._aligned(tensor.numel())
        assert aligned <= self.max_numel
        self._make_swap_space(aligned)
        buf = self.all_buffers[self.current_buffer_index]
        buf.insert_tensor(tensor, swap_path, aligned)

    def _make_swap_space(self, numel):
        if self.current_buffer_index == -1:
            self._allocate_buffer()
            return
        if not self.all_buffers[self.current_buffer_index].has_space(numel):
            if self.free_buffer_index:
                self._flush_ready_buffers()
            else:
                self._flush_buffers_until_complete()
            self._allocate_buffer()

    def _allocate_buffer(self):
        assert self.current_buffer_index == -1 and self.free_buffer_index
        self.current_buffer_index = self.free_buffer_index.pop()

    def _flush_ready_buffers(self):
        if self.current_buffer_index != -1:
            self.ready_buffer_index.append(self.current_buffer_index)
            self.current_buffer_index = -1
        self._swap_out_ready_buffers()

    def _flush_buffers_until_complete(self):
        self._flush_ready_buffers()
        if self.num_pending_swaps:
            self._wait_for_swap_complete()

    def _swap_out_ready_buffers(self):
        for i in self.ready_buffer_index:
            b = self.all_buffers[i]
            for t, p in zip(b.get_swap_tensors(), b.get_swap_paths()):
                self.aio_handle.async_pwrite(t, p)
                self.num_pending_swaps += 1
                self.num_elements_swapped += t.numel()
        self.swapping_buffer_index += self.ready_buffer_index
        self.ready_buffer_index = []

    def _wait_for_swap_complete(self):
        self.aio_handle.wait()
        self.num_pending_swaps = 0
        for i in self.swapping_buffer_index:
            self.all_buffers[i].reset()
        self.free_buffer_index += self.swapping_buffer_index
        self.swapping_buffer_index = []
