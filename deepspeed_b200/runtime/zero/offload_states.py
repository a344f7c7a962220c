"""Move selected ZeRO states between HBM and host at run time (reference ``runtime/zero/offload_states.py`` and
``engine.offload_states / reload_states``): frees HBM between training phases (e.g. RLHF generation)."""
from typing import Container, Optional

import torch

from .offload_config import OffloadDeviceEnum, OffloadStateTypeEnum


def _move(t: torch.Tensor, device, pin_memory=False, non_blocking=False):
    if t is None or t.device.type == torch.device(device).type:
        return t
    if torch.device(device).type == "cpu":
        dst = torch.empty(t.shape, dtype=t.dtype, device="cpu", pin_memory=pin_memory and torch.cuda.is_available())
        dst.copy_(t, non_blocking=non_blocking)
        return dst
    return t.to(device, non_blocking=non_blocking)


def offload_optimizer_states(zo, include: Optional[Container[OffloadStateTypeEnum]] = None,
                             device=OffloadDeviceEnum.cpu, pin_memory=True, non_blocking=False):
    """``zo``: ZeroShardedOptimizer.  Swaps the tensor *storage* in place so every view kept by the optimizer
    stays valid (``Tensor.data`` assignment)."""
    dev = "cpu" if str(getattr(device, "value", device)) == "cpu" else str(device)
    want = lambda k: include is None or k in include
    moved = getattr(zo, "_offloaded_states", {})
    if want(OffloadStateTypeEnum.optim_states):
        for k, t in zo.flat_opt.state_tensors().items():
            if torch.is_tensor(t):
                moved[f"optim:{k}"] = t.device
                t.data = _move(t.data, dev, pin_memory, non_blocking)
    if want(OffloadStateTypeEnum.hp_params) and torch.is_tensor(zo.master):  # (an NVMe-resident master is already off-device)
        moved["hp_params"] = zo.master.device
        zo.master.data = _move(zo.master.data, dev, pin_memory, non_blocking)
    if want(OffloadStateTypeEnum.lp_grads) or want(OffloadStateTypeEnum.contiguous_grad_buffer):
        if zo.grad_arena is not None:
            moved["grads"] = zo.grad_arena.device
            zo.grad_arena.data = _move(zo.grad_arena.data, dev, pin_memory, non_blocking)
    if want(OffloadStateTypeEnum.lp_params) and torch.is_tensor(getattr(zo, "lp_arena", None)) and zo.stage == 3:
        moved["lp_params"] = zo.lp_arena.device
        zo.lp_arena.data = _move(zo.lp_arena.data, dev, pin_memory, non_blocking)
    zo._offloaded_states = moved
    if not non_blocking and torch.cuda.is_available():
        torch.cuda.synchronize()
        torch.cuda.empty_cache()


def reload_optimizer_states(zo, non_blocking=False):
    moved = getattr(zo, "_offloaded_states", {})
    for key, dev in list(moved.items()):
        if key.startswith("optim:"):
            t = zo.flat_opt.state_tensors()[key[6:]]
        elif key == "hp_params":
            t = zo.master
        elif key == "grads":
            t = zo.grad_arena
        else:
            t = zo.lp_arena
        t.data = _move(t.data, dev, False, non_blocking)
    zo._offloaded_states = {}
    if not non_blocking and torch.cuda.is_available():
        torch.cuda.synchronize()


offload_states = offload_optimizer_states
reload_states = reload_optimizer_states


def offload_adam_states(optimizer, device, pin_memory: bool = False, non_blocking: bool = False):
    """Torch-optimizer flavour (reference ``offload_states.py:19``): moves ``exp_avg`` / ``exp_avg_sq``."""
    from deepspeed_b200.runtime.utils import offload_adam_states as _impl
    return _impl(optimizer, device, pin_memory=pin_memory, non_blocking=non_blocking)


def reload_adam_states(optimizer, device, non_blocking: bool = False):
    from deepspeed_b200.runtime.utils import reload_adam_states as _impl
    return _impl(optimizer, device, non_blocking=non_blocking)


def get_state_devices(model, state: OffloadStateTypeEnum):
    """Devices currently holding ``state`` for a ZeRO engine (reference ``offload_states.py:51``).  ``model`` is the
    engine (or anything with ``.optimizer`` being the sharded optimizer)."""
    zo = getattr(model, "optimizer", model)
    if state == OffloadStateTypeEnum.hp_params:
        return {zo.master.device} if getattr(zo, "master", None) is not None else set()
    if state == OffloadStateTypeEnum.lp_params:
        arena = getattr(zo, "lp_arena", None)
        return {arena.device} if arena is not None else {p.device for p in model.parameters()}
    if state in (OffloadStateTypeEnum.lp_grads, OffloadStateTypeEnum.contiguous_grad_buffer):
        return {zo.grad_arena.device} if getattr(zo, "grad_arena", None) is not None else set()
    if state == OffloadStateTypeEnum.optim_states:
        return {t.device for t in zo.flat_opt.state_tensors().values() if torch.is_tensor(t)}
    raise ValueError(f"unknown offload state {state}")
