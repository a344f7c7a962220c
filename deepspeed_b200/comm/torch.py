"""Object-style backend over ``torch.distributed`` (reference ``comm/torch.py:TorchBackend``).

``deepspeed_b200.comm`` is function-style (one module-level wrapper per collective, instrumented by the comms logger);
this class offers the same operations as methods for code that holds a backend object (``dist.cdb``) and owns the
global "comm off" debug switches.
"""
import os
from datetime import timedelta

import torch
import torch.distributed as td

from .backend import Backend
from .reduce_op import ReduceOp, to_torch

DS_COMM_ALL_GATHER_OFF = False
DS_COMM_REDUCE_SCATTER_OFF = False
DS_COMM_BROADCAST_OFF = False
DS_COMM_ALL_REDUCE_OFF = False
DS_COMM_REDUCE_OFF = False


def _switch(name, flag):
    globals()[f"DS_COMM_{name}_OFF"] = flag
    os.environ[f"DSB200_COMM_{name}_OFF"] = "1" if flag else "0"  # the function-style wrappers read the env switch


def all_gather_comm_off(flag=False):
    _switch("ALL_GATHER", flag)


def reduce_scatter_comm_off(flag=False):
    _switch("REDUCE_SCATTER", flag)


def broadcast_comm_off(flag=False):
    _switch("BROADCAST", flag)


def all_reduce_comm_off(flag=False):
    _switch("ALL_REDUCE", flag)


def reduce_comm_off(flag=False):
    _switch("REDUCE", flag)


def backward_comm_off(flag=False):
    """Turn off every collective the backward pass issues (compute-only profiling)."""
    all_gather_comm_off(flag)
    reduce_scatter_comm_off(flag)


def has_coalescing_manager():
    return hasattr(td.distributed_c10d, "_coalescing_manager")


def has_all_reduce_coalesced():
    return hasattr(td, "all_reduce_coalesced")


def build_shm_op():
    """Host shared-memory collectives (csrc/cpu/shm_comm.cpp) or None."""
    try:
        from . import shm
        return shm
    except Exception:
        return None


class Noop:

    def wait(self):
        return None


class TorchBackend(Backend):

    def __init__(self, backend, timeout=timedelta(minutes=30), init_method=None, rank=-1, world_size=-1, name="torch"):
        super().__init__(name=name)
        self.shm_comm_op = build_shm_op()
        self.using_mpi = False
        if not td.is_initialized():
            td.init_process_group(backend, timeout=timeout, init_method=init_method, rank=rank, world_size=world_size)
        self.initialized = True
        self.world_group = td.group.WORLD
        self.world_size, self.world_rank = td.get_world_size(), td.get_rank()

    # capability probes
    def has_all_gather_into_tensor(self):
        return hasattr(td, "all_gather_into_tensor")

    def has_reduce_scatter_tensor(self):
        return hasattr(td, "reduce_scatter_tensor")

    def get_all_gather_function(self):
        return td.all_gather_into_tensor

    def get_reduce_scatter_function(self):
        return td.reduce_scatter_tensor

    @staticmethod
    def _reduce_op(op):
        return to_torch(op) if isinstance(op, ReduceOp) else op

    # collectives
    def all_reduce(self, tensor, op=td.ReduceOp.SUM, group=None, async_op=False):
        if DS_COMM_ALL_REDUCE_OFF:
            return Noop()
        return td.all_reduce(tensor, self._reduce_op(op), group, async_op)

    def inference_all_reduce(self, tensor, op=td.ReduceOp.SUM, group=None):
        from . import comm
        return comm.inference_all_reduce(tensor, op=op, group=group)

    def all_reduce_coalesced(self, tensors, op=td.ReduceOp.SUM, group=None, async_op=False):
        from . import comm
        return comm.all_reduce_coalesced(tensors, op=op, group=group, async_op=async_op)

    def reduce(self, tensor, dst, op=td.ReduceOp.SUM, group=None, async_op=False):
        if DS_COMM_REDUCE_OFF:
            return Noop()
        return td.reduce(tensor, dst, self._reduce_op(op), group, async_op)

    def reduce_scatter(self, output, input_list, op=td.ReduceOp.SUM, group=None, async_op=False):
        if DS_COMM_REDUCE_SCATTER_OFF:
            return Noop()
        return td.reduce_scatter(output, input_list, self._reduce_op(op), group, async_op)

    def reduce_scatter_tensor(self, output_tensor, input_tensor, op=td.ReduceOp.SUM, group=None, async_op=False):
        if DS_COMM_REDUCE_SCATTER_OFF:
            return Noop()
        return td.reduce_scatter_tensor(output_tensor, input_tensor, self._reduce_op(op), group, async_op)

    def broadcast(self, tensor, src, group=None, async_op=False):
        if DS_COMM_BROADCAST_OFF:
            return Noop()
        return td.broadcast(tensor, src, group, async_op)

    def broadcast_object_list(self, object_list, src, group=None, device=None):
        return td.broadcast_object_list(object_list, src, group, device)

    def all_gather(self, tensor_list, tensor, group=None, async_op=False):
        if DS_COMM_ALL_GATHER_OFF:
            return Noop()
        return td.all_gather(tensor_list, tensor, group, async_op)

    def all_gather_into_tensor(self, output_tensor, input_tensor, group=None, async_op=False):
        if DS_COMM_ALL_GATHER_OFF:
            return Noop()
        return td.all_gather_into_tensor(output_tensor, input_tensor, group, async_op)

    all_gather_base = all_gather_into_tensor

    def all_gather_coalesced(self, output_tensors, input_tensors, group=None, async_op=False):
        from . import comm
        return comm.all_gather_coalesced(output_tensors, input_tensors, group=group, async_op=async_op)

    def all_to_all_single(self, output, input, output_split_sizes=None, input_split_sizes=None, group=None, async_op=False):
        from . import comm
        return comm.all_to_all_single(output, input, output_split_sizes, input_split_sizes, group=group, async_op=async_op)

    def all_to_all(self, output_tensor_list, input_tensor_list, group=None, async_op=False):
        from . import comm
        return comm.all_to_all(output_tensor_list, input_tensor_list, group=group, async_op=async_op)

    def send(self, tensor, dst, group=None, tag=0):
        return td.send(tensor, dst, group, tag)

    def recv(self, tensor, src=None, group=None, tag=0):
        return td.recv(tensor, src, group, tag)

    def isend(self, tensor, dst, group=None, tag=0):
        return td.isend(tensor, dst, group, tag)

    def irecv(self, tensor, src=None, group=None, tag=0):
        return td.irecv(tensor, src, group, tag)

    def gather(self, tensor, gather_list=None, dst=0, group=None, async_op=False):
        return td.gather(tensor, gather_list, dst, group, async_op)

    def scatter(self, tensor, scatter_list=None, src=0, group=None, async_op=False):
        return td.scatter(tensor, scatter_list, src, group, async_op)

    def barrier(self, group=None, async_op=False, device_ids=None):
        return td.barrier(group=group, async_op=async_op, device_ids=device_ids)

    def monitored_barrier(self, group=None, timeout=None, wait_all_ranks=False):
        return td.monitored_barrier(group=group, timeout=timeout, wait_all_ranks=wait_all_ranks)

    # bookkeeping
    def get_rank(self, group=None):
        return td.get_rank(group)

    def get_world_size(self, group=None):
        return td.get_world_size(group)

    def is_initialized(self):
        return td.is_initialized()

    def get_backend(self, group=None):
        return td.get_backend(group)

    def new_group(self, ranks):
        return td.new_group(ranks)

    def get_global_rank(self, group, group_rank):
        return td.get_global_rank(group, group_rank)

    def get_world_group(self):
        return td.group.WORLD

    def destroy_process_group(self, group=None):
        return td.destroy_process_group(group)

    def init_device_mesh(self, mesh_shape, mesh_dim_names):
        from torch.distributed.device_mesh import init_device_mesh
        dev = "cuda" if torch.cuda.is_available() else "cpu"
        return init_device_mesh(dev, mesh_shape, mesh_dim_names=mesh_dim_names)


def disable_compiler_collective(func):
    """Decorator kept for parity (reference ``comm/torch.py:23``): collectives are never traced here (no ``torch.compile`` on
    the hot path), so the function is returned unchanged."""
    return func


def get_coalescing_manager(group, device, reqs, async_op):
    """Context manager that batches the collectives issued inside it into one NCCL group call."""
    import torch
    return torch.distributed.distributed_c10d._coalescing_manager(group, device=device, async_ops=async_op)
