"""``zero.Init`` / ``GatheredParameters`` / external-parameter registry.

Parity target: reference ``runtime/zero/partition_parameters.py`` (``Init :824``,
``GatheredParameters :2121``, ``register_external_parameter :128``, ``ZeroParamStatus :214``).

Mechanism (different from the reference, which rewrites ``__init__`` of every ``nn.Module``
subclass, ``partition_parameters.py:374-571``): PyTorch's *global module registration hooks* tell us
when a fully-constructed sub-module is attached to its parent; at that moment every parameter of
the sub-module that is still whole is broadcast from rank 0, sliced to this rank's ``1/world``
piece (``p.ds_tensor``) and its full storage is dropped.  Parameters of the root module are
handled when the context exits.  ``nn.Module.apply`` is wrapped for the duration of the context
so weight-init callbacks (HF ``post_init``) see gathered parameters.  When the engine later
builds its unit plan it consumes the slices (``materialize_full``) unit by unit, so the full model
never resides on one device.
"""
import contextlib
import weakref
from enum import Enum
from typing import Iterable, List, Optional

import torch
from torch import nn
from torch.nn.modules import module as _nn_module

from deepspeed_b200 import comm as dist
from deepspeed_b200.accelerator import get_accelerator
from deepspeed_b200.utils.logging import logger
from .gather_handles import (AllGatherCoalescedHandle, AllGatherHandle, AllReduceCoalescedHandle, CUDAQuantizer,  # noqa: F401
                             MultipleAllGatherHandles, NoGatherCoalescedHandle, NoGatherHandle, QuantizationInfo)

_init_stack: List["Init"] = []
zero_init_context = 0


class ZeroParamType(Enum):
    NORMAL = 1
    PARTITIONED = 2
    REMOTE = 3


class ZeroParamStatus(Enum):
    AVAILABLE = 1
    NOT_AVAILABLE = 2
    INFLIGHT = 3


def is_zero_param(p) -> bool:
    return hasattr(p, "ds_numel")


def _world(group):
    return dist.get_world_size(group) if dist.is_initialized() else 1


def _rank(group):
    return dist.get_rank(group) if dist.is_initialized() else 0


def _slice_len(numel, world):
    return (numel + world - 1) // world


@torch.no_grad()
def partition_param(p: nn.Parameter, group=None, device=None, dtype=None, pin=False, broadcast=True):
    """Convert a whole parameter into a ZeRO parameter holding only this rank's slice."""
    if is_zero_param(p) and getattr(p, "ds_tensor", None) is not None:
        return
    world, rank = _world(group), _rank(group)
    if p.device.type == "meta":
        raise RuntimeError("zero.Init cannot shard a meta-device parameter; construct with a real device")
    full = p.data
    if dtype is not None and full.is_floating_point() and full.dtype != dtype:
        full = full.to(dtype)
    if broadcast and world > 1:
        src = dist.get_global_rank(group, 0) if group is not None else 0
        if full.device.type == "cpu" and dist.get_backend(group) == "nccl":
            tmp = full.cuda()
            dist.broadcast(tmp, src=src, group=group)
            full = tmp.cpu()
        else:
            full = full.contiguous()
            dist.broadcast(full, src=src, group=group)
    n = full.numel()
    sl = _slice_len(n, world)
    flat = full.reshape(-1)
    piece = torch.zeros(sl, dtype=full.dtype, device=device or full.device)
    lo, hi = rank * sl, min((rank + 1) * sl, n)
    if hi > lo:
        piece[:hi - lo].copy_(flat[lo:hi])
    if pin and piece.device.type == "cpu" and torch.cuda.is_available():
        piece = piece.pin_memory()
    p.ds_shape = full.shape
    p.ds_numel = n
    p.ds_tensor = piece
    p.ds_group = group
    p.ds_status = ZeroParamStatus.NOT_AVAILABLE
    p.ds_id = getattr(p, "ds_id", _next_id())
    p.data = torch.empty(0, dtype=full.dtype, device=full.device)
    _attach_methods(p)


_id_counter = [0]


def _next_id():
    _id_counter[0] += 1
    return _id_counter[0]


@torch.no_grad()
def materialize_full(p: nn.Parameter, device=None) -> torch.Tensor:
    """Return the full tensor of a ZeRO parameter (all-gather of the per-rank slices)."""
    zo = _owner(p)
    if zo is not None:  # adopted by a sharded optimizer: its unit arena is the source of truth
        return zo.get_full_lp_param(p)
    piece = getattr(p, "ds_tensor", None)
    if piece is None:
        return p.data
    group = getattr(p, "ds_group", None)
    world = _world(group)
    dev = device or (get_accelerator().current_device_name() if torch.cuda.is_available() else piece.device)
    local = piece.to(dev)
    if world > 1:
        out = torch.empty(local.numel() * world, dtype=local.dtype, device=dev)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
    else:
        out = local
    return out[:p.ds_numel].view(p.ds_shape)


def _owner(p):
    ref = getattr(p, "_ds_zero", None)
    return ref() if ref is not None else None


def _attach_methods(p):
    """Per-parameter convenience API mirroring the reference (``param.all_gather()``, ``.partition()``)."""

    def all_gather(param_list=None, async_op=False, hierarchy=0):
        qs = list(param_list or [p])
        for q in qs:  # adopted by a sharded optimizer: gather the owning unit outside its rotating pool
            zo = _owner(q)
            if zo is not None and not getattr(q, "_ds_user_gathered", False):
                zo.gather_param_temp(q)
                q._ds_user_gathered = True
        todo = [q for q in qs if _owner(q) is None and q.data.numel() == 0 and getattr(q, "ds_tensor", None) is not None]
        if async_op:
            return all_gather_coalesced(todo)
        for q in todo:
            q.data = materialize_full(q)
            q.ds_status = ZeroParamStatus.AVAILABLE

    def partition(param_list=None, hierarchy=0, has_been_updated=False):
        for q in (param_list or [p]):
            zo = _owner(q)
            if zo is not None:
                if has_been_updated and q.data.numel() == q.ds_numel:
                    zo.set_full_hp_param(q.data, q)
                if getattr(q, "_ds_user_gathered", False):
                    q._ds_user_gathered = False
                    zo.release_param_temp(q)
                continue
            if has_been_updated and q.data.numel():
                _write_back(q, q.data)
            if getattr(q, "ds_tensor", None) is not None:
                q.data = torch.empty(0, dtype=q.dtype, device=q.device)
                q.ds_status = ZeroParamStatus.NOT_AVAILABLE

    p.all_gather = all_gather
    p.partition = partition
    p.ds_summary = lambda: {
        "id": p.ds_id,
        "status": p.ds_status.name if isinstance(p.ds_status, Enum) else p.ds_status,
        "numel": p.numel(),
        "ds_numel": getattr(p, "ds_numel", p.numel()),
        "shape": tuple(p.shape),
        "ds_shape": tuple(getattr(p, "ds_shape", p.shape)),
        "requires_grad": p.requires_grad,
    }


@torch.no_grad()
def all_gather_coalesced(params, quantize=False):
    """Launch ONE asynchronous all-gather for ``params`` (same process group, same dtype per bucket) and return a handle;
    ``handle.wait()`` installs the full tensors (reference ``Init._all_gather_dtype`` / ``all_gather_coalesced``)."""
    params = list(params)
    for q in params:
        q.ds_status = ZeroParamStatus.INFLIGHT
    if not params:
        return MultipleAllGatherHandles([])
    by_key = {}
    for q in params:
        by_key.setdefault((q.ds_tensor.dtype, id(getattr(q, "ds_group", None))), []).append(q)
    handles = []
    for (_, _), bucket in by_key.items():
        group = getattr(bucket[0], "ds_group", None)
        world = _world(group)
        if world == 1:
            handles.append(NoGatherCoalescedHandle(bucket))
            continue
        dev = get_accelerator().current_device_name() if torch.cuda.is_available() else bucket[0].ds_tensor.device
        local = torch.cat([q.ds_tensor.reshape(-1).to(dev) for q in bucket])
        if quantize and local.numel() % 8 == 0:
            info = QuantizationInfo()
            info.backend = CUDAQuantizer()
            qv, scales = info.backend.quantize(local)
            info.partition_sz, info.world_size = local.numel(), world
            info.quantized_param = torch.empty(world * qv.numel(), dtype=qv.dtype, device=qv.device)
            info.scale_buffer = torch.empty(world * scales.numel(), dtype=scales.dtype, device=scales.device)
            work = dist.all_gather_into_tensor(info.quantized_param, qv.reshape(-1), group=group, async_op=True)
            info.quant_handle = dist.all_gather_into_tensor(info.scale_buffer, scales.reshape(-1), group=group, async_op=True)
            handles.append(AllGatherCoalescedHandle(work, bucket, None, world, quantization=info))
            continue
        flat = torch.empty(world * local.numel(), dtype=local.dtype, device=dev)
        work = dist.all_gather_into_tensor(flat, local, group=group, async_op=True)
        parts = [flat.narrow(0, r * local.numel(), local.numel()) for r in range(world)]
        handles.append(AllGatherCoalescedHandle(work, bucket, parts, world))
    return handles[0] if len(handles) == 1 else MultipleAllGatherHandles(handles)


@torch.no_grad()
def free_param(param) -> None:
    """Release the gathered storage of a ZeRO parameter (its slice stays in ``ds_tensor``); reference ``:282``."""
    assert not getattr(param, "ds_active_sub_modules", None), param.ds_summary()
    if param.data.is_cuda:
        param.data.record_stream(torch.cuda.current_stream())
    param.data = torch.empty(0, dtype=param.dtype, device=param.device)
    param.ds_status = ZeroParamStatus.NOT_AVAILABLE


def get_all_subclasses(cls, include_root=True):
    """Transitive subclasses of ``cls``."""
    found, stack = set(), [cls]
    while stack:
        for sub in stack.pop().__subclasses__():
            if sub not in found:
                found.add(sub)
                stack.append(sub)
    if include_root:
        found.add(cls)
    return found


def _local_device():
    import os
    return torch.device(get_accelerator().device_name(int(os.environ.get("LOCAL_RANK", 0)))) if torch.cuda.is_available() \
        else torch.device("cpu")


def zero_wrapper_for_fp_tensor_constructor(fn, target_fp_dtype):
    """Wrap ``torch.empty``-like constructors so float tensors are born on the local device in ``target_fp_dtype``
    (reference ``:235``; used while a model is constructed under ``zero.Init(dtype=...)``)."""

    def wrapped_fn(*args, **kwargs):
        if kwargs.get("device") is None:
            kwargs["device"] = _local_device()
        t = fn(*args, **kwargs)
        if t.is_floating_point():
            t.data = t.data.to(target_fp_dtype)
        return t

    return wrapped_fn


def get_new_tensor_fn_for_dtype(dtype):
    """Replacement for ``Tensor.new_tensor``-style class constructors: float results are cast to ``dtype``."""

    def new_tensor(cls, *args, **kwargs):
        t = torch.empty(0, device=_local_device()).new_empty(*(args or (0, )), **kwargs)
        return t.to(dtype) if t.is_floating_point() else t

    return new_tensor


def print_rank_0(message, debug=False, force=False):
    if (debug or force) and (not dist.is_initialized() or dist.get_rank() == 0):
        print(message)


def debug_rank0(msg):
    if not dist.is_initialized() or dist.get_rank() == 0:
        logger.debug(msg)


@torch.no_grad()
def _write_back(p, full):
    """Store ``full`` back into the per-rank slice of an Init-sharded parameter."""
    zo = _owner(p)
    if zo is not None:
        zo.set_full_hp_param(full, p)
        return
    piece = p.ds_tensor
    world, rank = _world(getattr(p, "ds_group", None)), _rank(getattr(p, "ds_group", None))
    sl = piece.numel()
    flat = full.reshape(-1)
    lo, hi = rank * sl, min((rank + 1) * sl, p.ds_numel)
    if hi > lo:
        piece[:hi - lo].copy_(flat[lo:hi].to(piece.device, piece.dtype))


class Init(contextlib.AbstractContextManager):
    """Construct a model with parameters sharded across the data-parallel group as they are created.

    Arguments follow the reference (``partition_parameters.py:833``): ``module`` (shard an existing
    module in place), ``data_parallel_group``, ``mem_efficient_linear``, ``remote_device``
    (``"cpu"``/``"nvme"`` keep slices on the host), ``pin_memory``, ``config_dict_or_path``,
    ``enabled``, ``dtype``, ``mpu``, ``zero_param_parallel_group``, ``zero_quantized_weights``.
    """

    def __init__(self, module=None, data_parallel_group=None, mem_efficient_linear=True, remote_device=None,
                 pin_memory=False, config_dict_or_path=None, config=None, enabled=True, dtype=None, mpu=None,
                 zero_param_parallel_group=None, zero_quantized_weights=False,
                 zero_quantized_nontrainable_weights=False, sequence_data_parallel_group=None, param_swapper=None):
        self.enabled = enabled
        self.group = data_parallel_group or sequence_data_parallel_group
        self.remote_device = remote_device
        self.pin_memory = pin_memory
        self.mem_efficient_linear = mem_efficient_linear
        cfg = config_dict_or_path if config_dict_or_path is not None else config
        self.dtype = dtype
        if cfg is not None and dtype is None:
            from deepspeed_b200.runtime.config import DeepSpeedConfig
            c = cfg if isinstance(cfg, DeepSpeedConfig) else DeepSpeedConfig(cfg, mpu)
            if c.fp16_enabled:
                self.dtype = torch.float16
            elif c.bfloat16_enabled:
                self.dtype = torch.bfloat16
            op = c.zero_config.offload_param
            if op is not None and self.remote_device is None and str(getattr(op.device, "value", op.device)) != "none":
                self.remote_device = str(getattr(op.device, "value", op.device))
                self.pin_memory = bool(op.pin_memory)
        self._handles = []
        self._seen_params = []
        self._orig_apply = None
        self._device_ctx = None
        if module is not None and enabled:
            if not dist.is_initialized():
                dist.init_distributed()
            self._shard_module(module)

    # ---- context protocol ----------------------------------------------------------------------
    def __enter__(self):
        global zero_init_context
        if not self.enabled:
            return self
        if not dist.is_initialized():
            dist.init_distributed()
        _init_stack.append(self)
        zero_init_context += 1
        self._handles.append(_nn_module.register_module_module_registration_hook(self._on_child_registered))
        self._handles.append(_nn_module.register_module_parameter_registration_hook(self._on_param_registered))
        self._orig_apply = nn.Module.apply
        init = self

        def gathered_apply(module, fn):
            # gather -> apply -> re-shard, module by module (reference: override_module_apply)
            for child in module.children():
                gathered_apply(child, fn)
            own = [p for p in module.parameters(recurse=False) if is_zero_param(p)]
            with GatheredParameters(own, modifier_rank=0, enabled=bool(own)):
                fn(module)
            return module

        nn.Module.apply = gathered_apply
        # construct directly on the accelerator so init kernels run on the GPU
        if torch.cuda.is_available():
            self._device_ctx = torch.device(get_accelerator().current_device_name())
            self._device_ctx.__enter__()
        return self

    def __exit__(self, exc_type, exc, tb):
        global zero_init_context
        if not self.enabled:
            return False
        if self._device_ctx is not None:
            self._device_ctx.__exit__(exc_type, exc, tb)
            self._device_ctx = None
        for h in self._handles:
            h.remove()
        self._handles.clear()
        nn.Module.apply = self._orig_apply
        if _init_stack and _init_stack[-1] is self:
            _init_stack.pop()
        zero_init_context -= 1
        if exc_type is None:
            # root-module parameters (never attached to a parent) are sharded now
            for ref in self._seen_params:
                p = ref()
                if p is not None and not is_zero_param(p):
                    self._shard_param(p)
        self._seen_params.clear()
        return False

    # ---- group facts (reference ``Init.get_partition_dp_group`` etc.) -----------------------------------------------
    def get_partition_dp_group(self, param=None):
        return getattr(param, "ds_group", None) if param is not None and hasattr(param, "ds_group") else self.group

    def get_partition_rank(self):
        return _rank(self.group)

    @property
    def num_partitions(self):
        return _world(self.group)

    def get_dp_process_group(self):
        return self.group

    # ---- hooks -----------------------------------------------------------------------------------
    def _on_param_registered(self, module, name, param):
        if param is not None:
            self._seen_params.append(weakref.ref(param))
        return None

    def _on_child_registered(self, module, name, submodule):
        if submodule is not None:
            self._shard_module(submodule)
        return None

    def _shard_module(self, m: nn.Module):
        for p in m.parameters(recurse=True):
            if not is_zero_param(p):
                self._shard_param(p)

    def _shard_param(self, p):
        if p.device.type == "meta":
            # ``torch.nn.utils.skip_init`` / ``device="meta"`` construction: nothing to shard yet. The parameter is picked
            # up once it has storage (``to_empty`` then attachment to the parent, or the sweep when the context exits).
            return
        dev = None
        if self.remote_device in ("cpu", "nvme"):
            dev = "cpu"
        if p.device.type == "cpu" and torch.cuda.is_available() and dev is None:
            p.data = p.data.to(get_accelerator().current_device_name())
        partition_param(p, self.group, device=dev, dtype=self.dtype, pin=self.pin_memory)


def shutdown_init_context():
    """Suspend an active ``zero.Init`` (``deepspeed.initialize`` calls this; reference :137)."""
    for ctx in list(_init_stack):
        for h in ctx._handles:
            h.remove()
        ctx._handles.clear()
        if ctx._orig_apply is not None:
            nn.Module.apply = ctx._orig_apply


def restore_init_context():
    pass


class GatheredParameters(contextlib.AbstractContextManager):
    """Temporarily gather ZeRO-sharded parameters (reference :2121).

    ``modifier_rank``: if not ``None`` the values as modified on that rank are broadcast and
    written back to every rank's shard (and fp32 master) on exit.
    """

    def __init__(self, params, modifier_rank: Optional[int] = None, fwd_module=None, enabled=True):
        self.enabled = enabled
        if params is None:
            params = []
        elif isinstance(params, nn.Parameter) or torch.is_tensor(params):
            params = [params]
        elif isinstance(params, nn.Module):
            params = list(params.parameters())
        self.params = [p for p in params if p is not None]
        self.modifier_rank = modifier_rank
        self._gathered = []

    def __enter__(self):
        if not self.enabled:
            return self
        for p in self.params:
            zo = _owner(p)
            if zo is not None:
                if zo.param_is_gathered(p):
                    continue
                zo.gather_param_temp(p)
                self._gathered.append((p, zo))
            elif getattr(p, "ds_tensor", None) is not None and p.data.numel() == 0:
                p.data = materialize_full(p).clone()
                p.ds_status = ZeroParamStatus.AVAILABLE
                self._gathered.append((p, None))
        return self

    def __exit__(self, *exc):
        if not self.enabled:
            return False
        for p, zo in self._gathered:
            if zo is not None:
                zo.release_param_temp(p, write_back_from=self.modifier_rank)
            else:
                if self.modifier_rank is not None:
                    grp = getattr(p, "ds_group", None)
                    if _world(grp) > 1:
                        src = dist.get_global_rank(grp, self.modifier_rank) if grp is not None else self.modifier_rank
                        dist.broadcast(p.data, src=src, group=grp)
                    _write_back(p, p.data)
                p.data = torch.empty(0, dtype=p.dtype, device=p.device)
                p.ds_status = ZeroParamStatus.NOT_AVAILABLE
        # params that were already resident but modified in place
        if self.modifier_rank is not None:
            for p in self.params:
                zo = _owner(p)
                if zo is not None and all(p is not q for q, _ in self._gathered):
                    zo.sync_param_from_full(p, src_rank=self.modifier_rank)
        self._gathered.clear()
        return False


# ---- external parameters: parameters used by a module that does not own them ----------------------
def register_external_parameter(module: nn.Module, parameter: nn.Parameter):
    """Declare that ``module.forward`` reads ``parameter`` owned elsewhere (reference :128).  The unit
    owning the parameter is fetched together with ``module``'s own unit."""
    if not isinstance(parameter, nn.Parameter):
        raise RuntimeError("Parameter is not a torch.nn.Parameter")
    if not hasattr(module, "_external_params"):
        module._external_params = {}
    key = id(parameter)
    module._external_params[key] = parameter
    zo = _owner(parameter)
    if zo is not None:
        zo.add_external_dependency(module, parameter)


def unregister_external_parameter(module: nn.Module, parameter: nn.Parameter):
    if hasattr(module, "_external_params"):
        module._external_params.pop(id(parameter), None)


# the reference's name for the machinery ``Init`` derives from (``partition_parameters.py:302``): here construction-time
# sharding is driven by torch's module/parameter registration hooks, so ``Init`` itself is that base
InsertPostInitMethodToModuleSubClasses = Init
