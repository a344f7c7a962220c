"""Safe accessors for sharded high-precision state.

Parity target: reference ``utils/tensor_fragment.py:132-299`` (``safe_get_full_fp32_param``,
``safe_get_full_grad``, ``safe_get_full_optimizer_state`` and their ``set`` / ``local`` variants).
The reference attaches a ``tensor_fragment`` object to every parameter; here the mapping lives in the
static unit plan (``runtime/zero/units.py param_fragments``) and the optimizer exposes gather /
scatter helpers on top of it, so these functions are thin dispatchers.  Every ``full`` accessor is a
collective: all ranks of the DP group must call it.
"""
import torch


def _zo(param):
    ref = getattr(param, "_ds_zero", None)
    return ref() if ref is not None else None


def safe_get_full_fp32_param(param):
    zo = _zo(param)
    if zo is None:
        return param.detach().float()
    return zo.get_full_hp_param(param)


def safe_set_full_fp32_param(param, value):
    zo = _zo(param)
    if zo is None:
        with torch.no_grad():
            param.copy_(value.to(param.dtype))
        return
    zo.set_full_hp_param(value, param)


def safe_get_full_optimizer_state(param, optim_state_key):
    zo = _zo(param)
    if zo is None:
        return None
    return zo.get_full_optimizer_state(param, optim_state_key)


def safe_set_full_optimizer_state(param, value, optim_state_key):
    zo = _zo(param)
    if zo is not None:
        zo.set_full_optimizer_state(value, param, optim_state_key)


def safe_get_full_grad(param):
    zo = _zo(param)
    if zo is None:
        return None if param.grad is None else param.grad.detach().float()
    return zo.get_full_hp_grad(param)


def safe_set_full_grad(param, value):
    zo = _zo(param)
    if zo is None:
        param.grad = value.to(param.dtype)
        return
    rt, s = zo.unit_of_param[id(param)], zo.slot_of_param[id(param)]
    zo._scatter_into_arena(zo.grad_arena, rt, s, value)


# ---- local (this rank's fragment) API: ZeRO-3 style ------------------------------------------------
def _local_view(zo, param, arena):
    from deepspeed_b200.runtime.zero.units import param_fragments
    if arena is None:
        return None
    rt, s = zo.unit_of_param[id(param)], zo.slot_of_param[id(param)]
    for (r, p0, a0, ln) in param_fragments(rt.u, s, zo.shard_world):
        if r == zo.shard_rank:
            return arena[a0:a0 + ln]
    return arena[0:0]


def safe_get_local_fp32_param(param):
    zo = _zo(param)
    if zo is None:
        return param.detach().float().view(-1)
    arena = zo.master if zo.master is not None else zo._lp_arena_as_flat()
    return _local_view(zo, param, arena).float()


def safe_set_local_fp32_param(param, value):
    zo = _zo(param)
    if zo is None:
        with torch.no_grad():
            param.view(-1).copy_(value)
        return
    v = _local_view(zo, param, zo.master if zo.master is not None else zo._lp_arena_as_flat())
    v.copy_(value.to(v.device, v.dtype).view(-1))
    if zo.master is not None and zo.lp_arena is not None:
        _local_view(zo, param, zo.lp_arena).copy_(value.view(-1))


def safe_get_local_grad(param):
    zo = _zo(param)
    if zo is None:
        return None if param.grad is None else param.grad.view(-1).float()
    v = _local_view(zo, param, zo.grad_arena)
    return None if v is None else v.float()


def safe_set_local_grad(param, value):
    zo = _zo(param)
    if zo is None:
        param.grad = value.view_as(param).to(param.dtype)
        return
    v = _local_view(zo, param, zo.grad_arena)
    if v is not None:
        v.copy_(value.to(v.device, v.dtype).view(-1))


def safe_get_local_optimizer_state(param, optim_state_key):
    zo = _zo(param)
    if zo is None:
        return None
    st = zo.flat_opt.state_tensors()
    if optim_state_key not in st:
        return None
    return _local_view(zo, param, st[optim_state_key]).float()


def safe_set_local_optimizer_state(param, value, optim_state_key):
    zo = _zo(param)
    if zo is None:
        return
    st = zo.flat_opt.state_tensors()
    v = _local_view(zo, param, st[optim_state_key])
    v.copy_(value.to(v.device, v.dtype).view(-1))


def get_hp_fragment_mapping(lp_param, lp_start=None, flat_hp_partition=None, gradient_dict=None, offload_gradient_dict=None,
                            use_offload=False, param_group_index=0, partition_start=None, partition_size=None):
    """Where does ``lp_param`` live in this rank's flat fp32 partition?  Returns a list of
    ``(param_start, arena_start, length)`` triples (reference ``tensor_fragment.py:312`` returns one ``tensor_fragment``
    because its partitions are per param group; here a parameter may straddle two ranks' shards of its unit)."""
    from deepspeed_b200.runtime.zero.units import param_fragments
    zo = _zo(lp_param)
    if zo is None:
        return []
    rt, slot = zo.unit_of_param[id(lp_param)], zo.slot_of_param[id(lp_param)]
    return [(p0, a0, ln) for (r, p0, a0, ln) in param_fragments(rt.u, slot, zo.shard_world) if r == zo.shard_rank]


# ---- reference data types (``utils/tensor_fragment.py:13-40``) -----------------------------------------------------------
from dataclasses import dataclass as _dataclass  # noqa: E402
from typing import Dict as _Dict  # noqa: E402


@_dataclass
class fragment_address:
    numel: int
    start: int


@_dataclass
class tensor_fragment:
    """One contiguous piece of a low-precision parameter inside this rank's flat high-precision partition."""
    lp_fragment: torch.Tensor
    lp_fragment_address: fragment_address
    hp_fragment: torch.Tensor
    hp_fragment_address: fragment_address
    gradient_dict: _Dict = None
    offload_gradient_dict: _Dict = None
    use_offload: bool = False
    param_group_index: int = 0
    optim_fragment: _Dict = None

    def update_hp(self):
        self.hp_fragment.data.copy_(self.lp_fragment.data)

    def update_lp(self):
        self.lp_fragment.data.copy_(self.hp_fragment.data)

    def get_optim_state_fragment(self, key):
        if self.optim_fragment is None or key not in self.optim_fragment:
            raise ValueError(f"{key} not found in optimizer state fragment")
        return self.optim_fragment[key]

    def set_optim_state_fragment(self, flat_hp_partition, optim_fragment):
        a = self.hp_fragment_address
        self.optim_fragment = {k: v.narrow(0, a.start, a.numel) for k, v in optim_fragment.items()
                               if torch.is_tensor(v) and v.dim() > 0 and v.numel() == flat_hp_partition.numel()}

    def get_hp_fragment_address(self):
        return self.hp_fragment_address

    def get_optim_state_keys(self):
        return list((self.optim_fragment or {}).keys())

    def get_hp_fragment(self, optim_state_key=None):
        return self.hp_fragment if optim_state_key is None else self.get_optim_state_fragment(optim_state_key)


def map_to_flat_opt_states(flat_hp_tensor, lp_tensors, optim_state, opt_keys):
    """Build flat optimizer-state buffers shaped like ``flat_hp_tensor`` out of the per-parameter states of
    ``lp_tensors`` (used when a torch optimizer's state must follow a flattened master copy)."""
    for key in opt_keys:
        buf = torch.zeros_like(flat_hp_tensor)
        offset = 0
        for lp in lp_tensors:
            st = optim_state.get(lp, {})
            if key in st and torch.is_tensor(st[key]) and st[key].numel() == lp.numel():
                buf.narrow(0, offset, lp.numel()).copy_(st[key].reshape(-1))
                st[key] = buf.narrow(0, offset, lp.numel()).view_as(lp)
            offset += lp.numel()
        optim_state.setdefault(flat_hp_tensor, {})[key] = buf


# --- method-style accessors (reference ``tensor_fragment.py:87-130``): the reference binds these on every low-precision
# parameter (``param.get_full_hp_param()``); ``link_hp_params``-style callers can bind them with ``types.MethodType``.
def get_full_hp_param(self, optim_state_key=None):
    """fp32 master weight (``optim_state_key=None``) or the named optimizer state, assembled over the DP group."""
    if optim_state_key is None:
        return safe_get_full_fp32_param(self)
    return safe_get_full_optimizer_state(self, optim_state_key)


def set_full_hp_param(self, value, optim_state_key=None):
    if optim_state_key is None:
        return safe_set_full_fp32_param(self, value)
    return safe_set_full_optimizer_state(self, value, optim_state_key)


def get_full_hp_grad(self):
    return safe_get_full_grad(self)


def set_full_hp_grad(self, value):
    return safe_set_full_grad(self, value)
