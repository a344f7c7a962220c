"""NVMe tier for optimizer state: file-backed flat arrays streamed through pinned windows.

Role parity: reference ``optimizer_utils.py:117 OptimizerSwapper``, ``partitioned_optimizer_swapper.py:27`` and
``pipelined_optimizer_swapper.py:52`` (swap-in of sub-group i+1 and swap-out of i-1 overlap the CPU step of i).
Because this framework's optimizer state is ONE flat array per state name (not a tensor per parameter), a
"sub-group" is simply an arena range: ``SwappedFlatState[s:e]`` hands out a pinned window of that range (read
from its file, or taken from an already-prefetched window) and writes the previous window back asynchronously.
"""
import os

import torch

from .utils import _pinned


class SwappedFlatState:
    """Tensor-like facade over ``<folder>/<name>.swp`` (``numel`` elements of ``dtype``)."""

    def __init__(self, name, numel, dtype, folder, aio_handle, window_elems, n_windows=3):
        self.name, self._numel, self.dtype = name, int(numel), dtype
        self.path = os.path.join(folder, f"{name}.swp")
        self.aio = aio_handle
        self.window_elems = int(min(window_elems, numel))
        self._bufs = [_pinned(self.window_elems, dtype) for _ in range(max(2, n_windows))]
        self._active = None      # (buf_idx, s, e) handed out, possibly dirty
        self._prefetched = None  # (buf_idx, s, e) read in flight / done
        self._writing = set()    # buf idxs with writes in flight
        self.device = torch.device("cpu")
        os.makedirs(folder, exist_ok=True)
        self._zero_fill()

    # ---- tensor-like surface used by the optimizer / checkpoint code
    def numel(self):
        return self._numel

    @property
    def shape(self):
        return torch.Size([self._numel])

    def _zero_fill(self):
        z = self._bufs[0]
        z.zero_()
        it = self.dtype.itemsize
        for s in range(0, self._numel, self.window_elems):
            e = min(s + self.window_elems, self._numel)
            self.aio.sync_pwrite(z[:e - s], self.path, s * it)

    def _free_buf(self):
        busy = {x[0] for x in (self._active, self._prefetched) if x is not None}
        for i in range(len(self._bufs)):
            if i not in busy and i not in self._writing:
                return i
        self._drain()
        for i in range(len(self._bufs)):
            if i not in busy:
                return i
        raise RuntimeError("no free swap window")

    def _drain(self):
        self.aio.wait()
        self._writing.clear()

    def flush(self, wait=True):
        """Write the active window back."""
        if self._active is not None:
            i, s, e = self._active
            self.aio.async_pwrite(self._bufs[i][:e - s], self.path, s * self.dtype.itemsize)
            self._writing.add(i)
            self._active = None
        if wait:
            self._drain()

    def prefetch(self, s, e):
        if self._prefetched is not None and self._prefetched[1:] == (s, e):
            return
        assert e - s <= self.window_elems
        i = self._free_buf()
        # async reads and writes share the handle's completion queue: a wait() completes both
        self.aio.async_pread(self._bufs[i][:e - s], self.path, s * self.dtype.itemsize)
        self._prefetched = (i, s, e)

    def __getitem__(self, sl):
        s, e, step = sl.indices(self._numel)
        assert step == 1
        if self._active is not None and self._active[1] <= s and e <= self._active[2]:
            i, s0, _ = self._active
            return self._bufs[i][s - s0:e - s0]
        if e - s > self.window_elems:
            # larger than one pinned window (initialisation, checkpoint / debug access to a whole shard): a range object
            # that reads / writes the file window by window; the optimizer step itself always asks for <= one window
            return _SwapRange(self, s, e)
        self.flush(wait=False)
        if self._prefetched is not None and self._prefetched[1:] == (s, e):
            i = self._prefetched[0]
            self._prefetched = None
            self._drain()
        else:
            if self._prefetched is not None:
                self._drain()
                self._prefetched = None
            i = self._free_buf()
            self.aio.sync_pread(self._bufs[i][:e - s], self.path, s * self.dtype.itemsize)
        self._active = (i, s, e)
        return self._bufs[i][:e - s]

    def read_range(self, a, b):
        """Elements [a, b) as an ordinary host tensor (window-by-window reads)."""
        self.flush()
        out = torch.empty(b - a, dtype=self.dtype)
        tmp = self._bufs[self._free_buf()]
        it = self.dtype.itemsize
        for s in range(a, b, self.window_elems):
            e = min(s + self.window_elems, b)
            self.aio.sync_pread(tmp[:e - s], self.path, s * it)
            out[s - a:e - a].copy_(tmp[:e - s])
        return out

    def write_range(self, a, src):
        """Overwrite elements [a, a + src.numel()) (window-by-window writes)."""
        self.flush()
        self._prefetched = None if self._prefetched is None or not (self._prefetched[1] < a + src.numel() and
                                                                    a < self._prefetched[2]) else self._drop_prefetch()
        src = src.reshape(-1)
        b = a + src.numel()
        assert 0 <= a and b <= self._numel
        tmp = self._bufs[self._free_buf()]
        it = self.dtype.itemsize
        for s in range(a, b, self.window_elems):
            e = min(s + self.window_elems, b)
            tmp[:e - s].copy_(src[s - a:e - a])
            self.aio.sync_pwrite(tmp[:e - s], self.path, s * it)

    def _drop_prefetch(self):
        self._drain()
        return None

    def detach(self):
        """Materialise the whole array (checkpoint save)."""
        return self.read_range(0, self._numel)

    def cpu(self):
        return self.detach()

    def clone(self):
        return self.detach()

    def copy_(self, src):
        """Overwrite the whole array (checkpoint load)."""
        self.flush()
        src = src.reshape(-1)
        assert src.numel() == self._numel
        tmp = self._bufs[self._free_buf()]
        it = self.dtype.itemsize
        for s in range(0, self._numel, self.window_elems):
            e = min(s + self.window_elems, self._numel)
            tmp[:e - s].copy_(src[s:e])
            self.aio.sync_pwrite(tmp[:e - s], self.path, s * it)
        return self

    def zero_(self):
        self._active = None
        self._zero_fill()
        return self


class _SwapRange:
    """``swapped[a:b]`` for a range larger than one pinned window: enough of the tensor surface for the code that
    initialises, checkpoints or inspects a whole shard (``copy_``, ``to``, ``float``, ``cpu``, ``clone``)."""

    def __init__(self, owner, a, b):
        self.owner, self.a, self.b = owner, a, b
        self.device, self.dtype = owner.device, owner.dtype

    def numel(self):
        return self.b - self.a

    @property
    def shape(self):
        return torch.Size([self.b - self.a])

    def copy_(self, src, non_blocking=False):
        src = src.reshape(-1)
        assert src.numel() == self.b - self.a
        self.owner.write_range(self.a, src.detach().to("cpu", self.dtype))
        return self

    def _read(self):
        return self.owner.read_range(self.a, self.b)

    def to(self, *args, **kw):
        return self._read().to(*args, **kw)

    def float(self):
        return self._read().float()

    def cpu(self):
        return self._read()

    def detach(self):
        return self._read()

    def clone(self):
        return self._read()

    def contiguous(self):
        return self._read()

    def view(self, *shape):
        return self._read().view(*shape)


class FlatStateSwapper:
    """Owns the aio handle + swap folder of one ZeRO optimizer instance and turns a flat optimizer's state
    dict into NVMe-backed arrays."""

    def __init__(self, swap_config, aio_config, base_folder, rank, dtype=torch.float32):
        from .aio_config import make_handle
        self.folder = os.path.join(base_folder, "zero_stage_3", "optimizer", f"rank{rank}")
        os.makedirs(self.folder, exist_ok=True)
        self.aio = make_handle(aio_config)
        self.window_elems = int(getattr(swap_config, "b200_swap_window", 1 << 26) or (1 << 26))
        self.n_windows = max(2, int(getattr(swap_config, "buffer_count", 4) or 4) // 2)
        self.pipeline = bool(getattr(swap_config, "pipeline_read", False) or getattr(swap_config, "pipeline_write", False)
                             or getattr(swap_config, "pipeline", False))
        self.dtype = dtype

    def wrap(self, flat_opt, numel):
        for n in flat_opt.state_names:
            flat_opt.state[n] = SwappedFlatState(n, numel, self.dtype, self.folder, self.aio, self.window_elems,
                                                 self.n_windows + (1 if self.pipeline else 0))
        return flat_opt

    def wrap_master(self, numel, dtype=torch.float32):
        """The fp32 master weights as a swapped array too (reference ``optimizer_utils.py:117``: the swapper owns the
        fp32 parameter AND its optimizer states) -- nothing of the optimizer then stays resident in host memory."""
        m = SwappedFlatState("fp32_master", numel, dtype, self.folder, self.aio, self.window_elems,
                             self.n_windows + (1 if self.pipeline else 0))
        self.extra = getattr(self, "extra", []) + [m]
        return m

    def _arrays(self, flat_opt):
        return [t for t in list(flat_opt.state.values()) + getattr(self, "extra", []) if isinstance(t, SwappedFlatState)]

    def prefetch(self, flat_opt, s, e):
        if not self.pipeline:
            return
        for t in self._arrays(flat_opt):
            t.prefetch(s, e)

    def flush(self, flat_opt, wait=True):
        for t in self._arrays(flat_opt):
            t.flush(wait=False)
        if wait:
            self.aio.wait()
            for t in self._arrays(flat_opt):
                t._writing.clear()


# ---- per-parameter swapper (reference ``optimizer_utils.py:21-480``) ----------------------------------------------------
# The flat swapper above is what this framework's own optimizers use.  ``OptimizerSwapper`` keeps the reference's
# per-parameter contract for client code that drives swapping itself with an ordinary ``torch.optim`` optimizer: the
# state tensors of one parameter live in ``<folder>/<param id>_<state>.tensor.swp`` between uses.
class FlattenedTensorSwapInfo:
    """A slice (``offset``, ``length`` elements) of a flattened tensor parked in ``path``."""

    def __init__(self, path, length, offset):
        self.path, self.length, self.offset = path, length, offset


class OptimizerStateSwapInfo:
    """Book-keeping of one parameter: its state tensors, their files, and gradient slices swapped out during backward."""

    def __init__(self, parameter, numel, base_folder):
        self.tensors, self.swap_paths = [], []
        self.param_id = OptimizerSwapper.parameter_id(parameter)
        self.swap_folder = base_folder
        self.swapped_gradients = {}    # offset -> FlattenedTensorSwapInfo
        self.unswapped_gradients = {}  # offset -> tensor kept in memory (too small / misaligned for aio)
        self.tensor_numel = numel
        self.tensor_dtype, self.tensor_device = parameter.dtype, parameter.device
        self.has_state_tensors = False
        self._add_tensors([parameter])

    def numel(self):
        return self.tensor_numel

    def has_gradients(self):
        return bool(self.swapped_gradients or self.unswapped_gradients)

    def _add_tensors(self, tensor_list):
        for t in tensor_list:
            self.tensors.append(t)
            self.swap_paths.append(os.path.join(self.swap_folder, f"{self.param_id}_{len(self.tensors) - 1}.tensor.swp"))

    def add_state_tensors(self, tensor_list):
        self.has_state_tensors = True
        self._add_tensors(tensor_list)

    def device(self):
        return self.tensor_device

    def dtype(self):
        return self.tensor_dtype

    def release_memory(self):
        for t in self.tensors:
            t.data = torch.empty(0, dtype=t.dtype, device=t.device)

    def get_or_create_gradient_paths(self, offsets, lengths):
        paths = []
        for off, n in zip(offsets, lengths):
            info = self.swapped_gradients.get(off)
            if info is None:
                info = FlattenedTensorSwapInfo(os.path.join(self.swap_folder, f"{self.param_id}_gradient_{off}_{n}.tensor.swp"),
                                               n, off)
                self.swapped_gradients[off] = info
            paths.append(info.path)
        return paths

    def set_swap_buffers(self, buffers):
        """Point every tensor at a window of its (pinned) buffer: where swap-in lands and compute happens."""
        for t, buf in zip(self.tensors, buffers):
            t.data = buf.narrow(0, 0, self.numel()).data

    def get_swap_gradient_buffers(self, swap_buffer):
        return [swap_buffer.narrow(0, g.offset, g.length) for g in self.swapped_gradients.values()]

    def get_swap_gradient_paths(self):
        return [g.path for g in self.swapped_gradients.values()]

    def get_unpinned_state_tensors(self):
        return [t for t in self.tensors if not t.is_pinned()]

    def read_unswapped_gradients(self, dest_buffer):
        n = 0
        for off, g in self.unswapped_gradients.items():
            dest_buffer.narrow(0, off, g.numel()).copy_(g)
            n += g.numel()
        return n

    def release_unswapped_gradients(self):
        self.unswapped_gradients = {}


SWAPPER_DEBUG_MODE = False
SWAP_OUT_GRADIENT_TIMER = "swap_out_gradient"


class OptimizerSwapper:
    """Per-parameter optimizer-state swapper over an aio handle."""

    @staticmethod
    def parameter_id(param):
        return getattr(param, "ds_id", id(param))

    def __init__(self, swap_config, aio_config, base_folder, optimizer, largest_numel, device, dtype, timers):
        from .aio_config import make_handle
        self.swap_config, self.aio_config = swap_config, aio_config
        self.swap_folder = os.path.join(base_folder, "optimizer", f"rank{_rank()}")
        os.makedirs(self.swap_folder, exist_ok=True)
        self.optimizer, self.device, self.dtype, self.timers = optimizer, device, dtype, timers
        self.aio_handle = make_handle(aio_config)
        self.swap_element_size = torch.tensor([], dtype=dtype).element_size()
        block = getattr(aio_config, "block_size", None) or (aio_config or {}).get("block_size", 1 << 20)
        threads = getattr(aio_config, "intra_op_parallelism", None) or (aio_config or {}).get("intra_op_parallelism", 1)
        self.min_aio_bytes = max(1 << 20, block)
        self.aligned_bytes = 4096 * max(1, threads)
        self.numel_alignment = self.aligned_bytes // self.swap_element_size
        self.largest_numel = self._io_aligned_numel(largest_numel)
        self.swap_params_info = {}
        self.timer_names = set()
        self.swappable_tensor_min_numel = self.min_aio_bytes // self.swap_element_size

    # ---- policy
    def swappable_tensor(self, param=None, numel=None):
        assert param is not None or numel is not None, "Either param or numel must be provided"
        n = param.numel() if param is not None else numel
        return self.min_aio_bytes <= n * self.swap_element_size

    def _io_aligned_numel(self, numel):
        return -(-numel // self.numel_alignment) * self.numel_alignment

    # ---- bookkeeping
    def _get_state_tensors(self, parameter):
        st = self.optimizer.state.get(parameter, {})
        return [v for v in st.values() if torch.is_tensor(v) and v.numel() == parameter.numel()]

    def _create_param_swap_info(self, parameter, numel):
        pid = self.parameter_id(parameter)
        assert pid not in self.swap_params_info
        info = OptimizerStateSwapInfo(parameter, numel, self.swap_folder)
        self.swap_params_info[pid] = info
        self._update_param_state_info(info, parameter)
        return info

    def _update_param_state_info(self, swap_info, parameter):
        if not swap_info.has_state_tensors:
            states = self._get_state_tensors(parameter)
            if states:
                swap_info.add_state_tensors(states)

    def _get_param_swap_info(self, parameter):
        info = self.swap_params_info.get(self.parameter_id(parameter))
        if info is not None:
            self._update_param_state_info(info, parameter)
        return info

    def purge_state(self):
        for info in self.swap_params_info.values():
            info.tensors = info.tensors[:1]
            info.swap_paths = info.swap_paths[:1]
            info.has_state_tensors = False

    # ---- state movement
    def _staging(self, numel):
        return _pinned(self._io_aligned_numel(numel), self.dtype)

    def swap_out_optimizer_state(self, parameter, async_swap=False):
        """Write the parameter (fp32 master) and its state tensors to their files and release the memory."""
        info = self._get_param_swap_info(parameter) or self._create_param_swap_info(parameter, parameter.numel())
        keep = []
        for t, path in zip(info.tensors, info.swap_paths):
            buf = self._staging(t.numel())
            buf[:t.numel()].copy_(t.detach().reshape(-1))
            self.aio_handle.async_pwrite(buf, path)
            keep.append(buf)
        self.aio_handle.wait()
        info.shapes = [tuple(t.shape) for t in info.tensors]
        info.release_memory()

    def swap_in_optimizer_state(self, parameter, async_parameter=None):
        """Bring the parameter and its states back (into fresh pinned windows)."""
        info = self._get_param_swap_info(parameter)
        if info is None:
            return
        bufs = [self._staging(info.numel()) for _ in info.tensors]
        for buf, path in zip(bufs, info.swap_paths):
            self.aio_handle.async_pread(buf, path)
        self.aio_handle.wait()
        for t, buf, shape in zip(info.tensors, bufs, getattr(info, "shapes", [None] * len(bufs))):
            t.data = buf[:info.numel()].view(shape if shape is not None else (info.numel(), ))
        if info.has_gradients():
            g = torch.zeros(info.numel(), dtype=self.dtype)
            self._retrieve_unswapped_grad_partitions(info, g)
            for sw in info.swapped_gradients.values():
                tmp = self._staging(sw.length)
                self.aio_handle.sync_pread(tmp, sw.path)
                g.narrow(0, sw.offset, sw.length).copy_(tmp[:sw.length])
            parameter.grad = g.view(parameter.shape)
            info.swapped_gradients = {}

    def swap_out_gradients(self, parameter, gradient_offsets, gradient_tensors):
        """Park gradient slices produced during backward: large aligned ones on NVMe, the rest in host memory."""
        info = self._get_param_swap_info(parameter) or self._create_param_swap_info(parameter, parameter.numel())
        for off, g in zip(gradient_offsets, gradient_tensors):
            if self.swappable_tensor(numel=g.numel()):
                path, = info.get_or_create_gradient_paths([off], [g.numel()])
                buf = self._staging(g.numel())
                buf[:g.numel()].copy_(g.detach().reshape(-1))
                self.aio_handle.sync_pwrite(buf, path)
            else:
                info.unswapped_gradients[off] = g.detach().reshape(-1).to("cpu", self.dtype).clone()

    def _retrieve_unswapped_grad_partitions(self, swap_info, dest_buffer):
        n = swap_info.read_unswapped_gradients(dest_buffer)
        swap_info.release_unswapped_gradients()
        return n

    # ---- hooks / timers
    def pre_backward(self):
        pass

    def post_backward(self):
        pass

    def init_timers(self):
        self.timer_names = set()

    def _start_timer(self, name):
        if self.timers:
            self.timers(name).start()
            self.timer_names.add(name)

    def _stop_timer(self, name):
        if self.timers:
            self.timers(name).stop()

    def log_timers(self):
        if self.timers and self.timer_names:
            self.timers.log(sorted(self.timer_names))


def _rank():
    from deepspeed_b200 import comm as dist
    return dist.get_rank() if dist.is_initialized() else 0
