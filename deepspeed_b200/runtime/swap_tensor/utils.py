"""Swap buffers and pools (reference ``runtime/swap_tensor/utils.py``: ``SwapBuffer :37``, ``SwapBufferPool
:96``, ``SwapBufferManager :180``)."""
import torch

MIN_AIO_BYTES = 1024**2
AIO_ALIGNED_BYTES = 1024


def swap_in_tensors(swap_handle, tensor_buffers, swap_paths):
    for buffer, path in zip(tensor_buffers, swap_paths):
        assert swap_handle.async_pread(buffer, path) >= 0


def swap_out_tensors(swap_handle, tensor_buffers, swap_paths):
    for buffer, path in zip(tensor_buffers, swap_paths):
        assert swap_handle.async_pwrite(buffer, path) >= 0


def get_sized_buffer(buffer, num_elems):
    assert num_elems <= buffer.numel(), f"num_elems {num_elems} > buffer {buffer.numel()}"
    return buffer.narrow(0, 0, num_elems) if num_elems < buffer.numel() else buffer


def get_sized_buffers(buffer_list, num_elems_list):
    return [get_sized_buffer(b, n) for b, n in zip(buffer_list, num_elems_list)]


def _pinned(numel, dtype):
    t = torch.empty(numel, dtype=dtype)
    return t.pin_memory() if torch.cuda.is_available() else t


class SwapBuffer:
    """A pinned staging buffer carved into tensors that each map to one swap file."""

    def __init__(self, buffer):
        self.buffer = buffer
        self.reset()

    def reset(self):
        self.offset = 0
        self.swap_tensors = {}
        self.compute_tensors = {}
        self.swap_paths = {}
        self.num_elem = 0

    def insert_tensor(self, tensor, swap_path, aligned_numel):
        swap_t, comp_t = self.allocate_tensor(swap_path, tensor.numel(), aligned_numel)
        comp_t.data.copy_(tensor.data.reshape(-1))
        return swap_t, comp_t

    def allocate_tensor(self, swap_path, numel, aligned_numel):
        assert self.has_space(aligned_numel)
        assert self.offset not in self.swap_tensors
        swap_t = self.buffer.narrow(0, self.offset, aligned_numel)
        comp_t = swap_t.narrow(0, 0, numel)
        self.swap_tensors[self.offset] = swap_t
        self.compute_tensors[self.offset] = comp_t
        self.swap_paths[self.offset] = swap_path
        self.offset += aligned_numel
        self.num_elem += numel
        return swap_t, comp_t

    def has_space(self, numel):
        return self.offset + numel <= self.buffer.numel()

    def get_swap_tensors(self):
        return list(self.swap_tensors.values())

    def get_swap_paths(self):
        return list(self.swap_paths.values())

    def get_compute_tensors(self):
        return list(self.compute_tensors.values())

    def get_num_elem(self):
        return self.num_elem

    def get_swap_tensor(self, offset):
        return self.swap_tensors.get(offset)

    def get_compute_tensor(self, offset):
        return self.compute_tensors.get(offset)

    def get_swap_path(self, offset):
        return self.swap_paths.get(offset)


class SwapBufferPool:

    def __init__(self, buffers):
        assert all(b.device.type == "cpu" for b in buffers)
        self.buffers = [SwapBuffer(b) for b in buffers]
        self.current_index = 0

    def reset(self):
        self.current_index = 0
        for b in self.buffers:
            b.reset()

    def _cur(self):
        return self.buffers[self.current_index]

    def allocate_tensor(self, numel, swap_path, aligned_numel):
        if self.has_space(aligned_numel):
            return self._cur().allocate_tensor(swap_path, numel, aligned_numel)
        return None, None

    def insert_tensor(self, tensor, swap_path, aligned_numel):
        if self.has_space(aligned_numel):
            return self._cur().insert_tensor(tensor, swap_path, aligned_numel)
        return None, None

    def get_swap_tensors(self):
        return [t for b in self._used() for t in b.get_swap_tensors()]

    def get_swap_paths(self):
        return [p for b in self._used() for p in b.get_swap_paths()]

    def get_compute_tensors(self):
        return [t for b in self._used() for t in b.get_compute_tensors()]

    def has_space(self, numel):
        if self._cur().has_space(numel):
            return True
        if self.current_index == len(self.buffers) - 1:
            return False
        self.current_index += 1
        return self._cur().has_space(numel)

    def swap_out(self, aio_handle, async_op=False):
        swap_out_tensors(aio_handle, self.get_swap_tensors(), self.get_swap_paths())
        if not async_op:
            assert len(self.get_swap_tensors()) == aio_handle.wait()

    def swap_in(self, aio_handle, async_op=False):
        swap_in_tensors(aio_handle, self.get_swap_tensors(), self.get_swap_paths())
        if not async_op:
            assert len(self.get_swap_tensors()) == aio_handle.wait()

    def _used(self):
        return self.buffers[:self.current_index + 1]


class SwapBufferManager:

    def __init__(self, num_elems, count, dtype):
        self.num_elems, self.count, self.dtype = num_elems, count, dtype
        self.all_buffers = [_pinned(num_elems, dtype) for _ in range(count)]
        self.free_buffer_index = list(range(count))
        self.used_buffer_index = {}
        self.gigabytes = self.all_buffers[0].element_size() * num_elems * count / 1024**3

    def allocate(self, num_elems, count, dtype):
        assert dtype == self.dtype and num_elems <= self.num_elems
        if count > len(self.free_buffer_index):
            return None
        idx = self.free_buffer_index[-count:]
        self.free_buffer_index = self.free_buffer_index[:-count]
        out = []
        for i in idx:
            t = self.all_buffers[i].narrow(0, 0, num_elems)
            out.append(t)
            self.used_buffer_index[id(t)] = i
        return out

    def allocate_all(self, num_elems, dtype):
        return self.allocate(num_elems, len(self.free_buffer_index), dtype)

    def free(self, buffers):
        for b in buffers:
            self.free_buffer_index.append(self.used_buffer_index.pop(id(b)))


def print_object(obj, name, exclude_list=()):
    """Log every attribute of ``obj`` as an aligned ``name ..... value`` table."""
    from deepspeed_b200.utils.logging import logger
    logger.info(f"{name}:")
    for arg in sorted(vars(obj)):
        if arg not in exclude_list:
            logger.info(f"  {arg} {'.' * max(1, 29 - len(arg))} {getattr(obj, arg)}")
