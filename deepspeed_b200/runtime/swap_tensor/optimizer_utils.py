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
        assert e - s <= self.window_elems, f"window {e - s} > {self.window_elems}; step in smaller pieces"
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

    def detach(self):
        """Materialise the whole array (checkpoint save)."""
        self.flush()
        out = torch.empty(self._numel, dtype=self.dtype)
        tmp = self._bufs[self._free_buf()]
        it = self.dtype.itemsize
        for s in range(0, self._numel, self.window_elems):
            e = min(s + self.window_elems, self._numel)
            self.aio.sync_pread(tmp[:e - s], self.path, s * it)
            out[s:e].copy_(tmp[:e - s])
        return out

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

    def prefetch(self, flat_opt, s, e):
        if not self.pipeline:
            return
        for t in flat_opt.state.values():
            if isinstance(t, SwappedFlatState):
                t.prefetch(s, e)

    def flush(self, flat_opt, wait=True):
        for t in flat_opt.state.values():
            if isinstance(t, SwappedFlatState):
                t.flush(wait=False)
        if wait:
            self.aio.wait()
            for t in flat_opt.state.values():
                if isinstance(t, SwappedFlatState):
                    t._writing.clear()
