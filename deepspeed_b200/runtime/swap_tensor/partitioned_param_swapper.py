"""NVMe tier for ZeRO-3 parameter shards (reference ``partitioned_param_swapper.py:36
AsyncPartitionedParameterSwapper``): each unit's low-precision shard has a swap file; ``swap_in`` brings a set
of shards into pinned buffers ahead of the all-gather, ``swap_out_and_release`` writes them back."""
import os
from enum import Enum

import torch

from .utils import _pinned


class PartitionedParamStatus(Enum):
    AVAILABLE = 1
    NOT_AVAILABLE = 2
    INFLIGHT = 3


class AsyncPartitionedParameterSwapper:

    def __init__(self, ds_config_or_offload, model_dtype, aio_config=None, base_folder=None, rank=0):
        from .aio_config import make_handle
        oc = ds_config_or_offload
        self.dtype = model_dtype
        self.folder = os.path.join(base_folder or str(getattr(oc, "nvme_path", "/tmp")), "zero_stage_3",
                                   f"{str(model_dtype).split('.')[-1]}params", f"rank{rank}")
        os.makedirs(self.folder, exist_ok=True)
        self.aio_read = make_handle(aio_config or {})
        self.aio_write = make_handle(aio_config or {})
        self.buffer_elems = int(getattr(oc, "buffer_size", 1 << 27) or (1 << 27))
        self.buffer_count = int(getattr(oc, "buffer_count", 5) or 5)
        self._pool = [_pinned(self.buffer_elems, model_dtype) for _ in range(self.buffer_count)]
        self._free = list(range(self.buffer_count))
        self._resident = {}   # id -> (buf idx, numel)
        self._inflight = set()
        self._status = {}
        self._numel = {}
        self.pending_writes = 0

    def _path(self, pid):
        return os.path.join(self.folder, f"{pid}_param.tensor.swp")

    def available_swap_in_buffers(self):
        return len(self._free)

    def swappable_tensor(self, param=None, numel=None):
        n = numel if numel is not None else param.numel()
        return n * self.dtype.itemsize >= 1024 and n <= self.buffer_elems

    def get_buffer(self, pid, numel):
        """Pinned tensor for ``pid`` (allocated on first use)."""
        if pid in self._resident:
            i, n = self._resident[pid]
            return self._pool[i][:n]
        if not self._free:
            raise RuntimeError("parameter swap buffers exhausted; raise offload_param.buffer_count")
        i = self._free.pop()
        self._resident[pid] = (i, numel)
        self._numel[pid] = numel
        return self._pool[i][:numel]

    def swap_out_and_release(self, pids, tensors=None, async_op=False, force_buffer_release=False):
        for k, pid in enumerate(pids):
            buf = self.get_buffer(pid, self._numel.get(pid, tensors[k].numel() if tensors else 0))
            if tensors is not None:
                buf.copy_(tensors[k].reshape(-1))
            self.aio_write.async_pwrite(buf, self._path(pid))
            self.pending_writes += 1
            self._status[pid] = PartitionedParamStatus.NOT_AVAILABLE
        if not async_op:
            self.synchronize_writes()
            for pid in pids:
                self._release(pid)

    def _release(self, pid):
        ent = self._resident.pop(pid, None)
        if ent is not None:
            self._free.append(ent[0])

    def synchronize_writes(self):
        if self.pending_writes:
            self.aio_write.wait()
            self.pending_writes = 0

    def swap_in(self, pids, async_op=True):
        out = []
        for pid in pids:
            buf = self.get_buffer(pid, self._numel[pid])
            self.aio_read.async_pread(buf, self._path(pid))
            self._inflight.add(pid)
            self._status[pid] = PartitionedParamStatus.INFLIGHT
            out.append(buf)
        if not async_op:
            self.synchronize_reads()
        return out

    def synchronize_reads(self):
        if self._inflight:
            self.aio_read.wait()
            for pid in self._inflight:
                self._status[pid] = PartitionedParamStatus.AVAILABLE
            self._inflight.clear()

    def release(self, pids):
        for pid in pids:
            self._release(pid)
            self._status[pid] = PartitionedParamStatus.NOT_AVAILABLE

    def status(self, pid):
        return self._status.get(pid, PartitionedParamStatus.NOT_AVAILABLE)

    # ---- parameter-object flavoured API of the reference (``partitioned_param_swapper.py:147-418``) --------------------
    # Ids are ``param.ds_id`` (or a unit id): the methods above take ids, these take the parameter objects.
    @staticmethod
    def _pid(param):
        return getattr(param, "ds_id", param)

    def get_path(self, param, must_exist=False):
        path = self._path(self._pid(param))
        assert not must_exist or os.path.exists(path), f"Path for param id {self._pid(param)} does not exist"
        return path

    def remove_partition_and_release_buffers(self, params):
        """Give the pinned buffers of ``params`` back to the pool (their data stays on NVMe)."""
        self.release([self._pid(p) for p in params])

    def swap_into_buffer(self, param, dest_buffer):
        """Synchronously read the shard of ``param`` into ``dest_buffer`` (through a pool buffer when the destination is
        not pinned / not a whole number of I/O blocks)."""
        pid = self._pid(param)
        n = self._numel.get(pid, dest_buffer.numel())
        pinned_ok = dest_buffer.is_pinned() and dest_buffer.is_contiguous()
        if pinned_ok:
            self.aio_read.sync_pread(dest_buffer[:n], self._path(pid))
        else:
            buf = self.swap_in([pid], async_op=False)[0]
            dest_buffer.reshape(-1)[:n].copy_(buf[:n])
            self._release(pid)
        self._status[pid] = PartitionedParamStatus.AVAILABLE

    def reserve_available_buffers(self):
        """Take every currently free pool buffer (used as scratch by the optimizer swapper); returns the tensors."""
        self._reserved = list(self._free)
        self._free = []
        return [self._pool[i] for i in self._reserved]

    def release_reserved_buffers(self):
        self._free.extend(getattr(self, "_reserved", []))
        self._reserved = []

    def reserve_partitioned_swap_space(self, partition_num_elems):
        """One pinned staging area large enough for all listed shards (each rounded to the I/O alignment)."""
        align = max(1, 4096 // self.dtype.itemsize)
        self._stage_offsets, pos = [], 0
        for n in partition_num_elems:
            self._stage_offsets.append((pos, n))
            pos += -(-n // align) * align
        self.partitioned_swap_buffer = _pinned(pos, self.dtype)

    def swap_out_partitioned_params(self, dst_fp16_params, src_fp32_params):
        """Cast the updated fp32 shards into the staging area and write each to its parameter's swap file."""
        assert getattr(self, "partitioned_swap_buffer", None) is not None, "call reserve_partitioned_swap_space first"
        assert len(dst_fp16_params) == len(src_fp32_params)
        self.synchronize_writes()
        for (pos, n), p, src in zip(self._stage_offsets, dst_fp16_params, src_fp32_params):
            pid = self._pid(p)
            stage = self.partitioned_swap_buffer[pos:pos + src.numel()]
            stage.copy_(src.reshape(-1))
            self._numel[pid] = src.numel()
            self.aio_write.async_pwrite(stage, self._path(pid))
            self.pending_writes += 1
            self._status[pid] = PartitionedParamStatus.NOT_AVAILABLE
        self.synchronize_writes()


def print_rank_0(message, debug=False, force=False):
    from deepspeed_b200 import comm as dist
    if (debug or force) and (not dist.is_initialized() or dist.get_rank() == 0):
        print(message)
