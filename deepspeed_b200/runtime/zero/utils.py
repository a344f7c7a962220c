"""ZeRO helper predicates (reference: ``runtime/zero/utils.py``)."""
import torch

from deepspeed_b200 import comm as dist


def _supported():
    from deepspeed_b200.ops.adam import DeepSpeedCPUAdam, FusedAdam
    from deepspeed_b200.ops.adagrad import DeepSpeedCPUAdagrad
    from deepspeed_b200.ops.lion import DeepSpeedCPULion, FusedLion
    return [torch.optim.Adam, torch.optim.AdamW, torch.optim.SGD, torch.optim.Adagrad, FusedAdam, DeepSpeedCPUAdam,
            DeepSpeedCPUAdagrad, DeepSpeedCPULion, FusedLion]


def is_zero_supported_optimizer(optimizer):
    return type(optimizer) in _supported()


def assert_ints_same_as_other_ranks(ints, group=None):
    """Cross-rank consistency check used in safe mode (reference :80)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.tensor(list(ints), dtype=torch.int64, device=dev)
    hi, lo = t.clone(), t.clone()
    dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=group)
    dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=group)
    if not (torch.equal(hi, t) and torch.equal(lo, t)):
        raise RuntimeError(f"disagreement between rank {dist.get_rank()} and other ranks: {list(ints)}")


def get_lst_from_rank0(lst, group=None):
    obj = [list(lst)]
    dist.broadcast_object_list(obj, src=0, group=group)
    return obj[0]


# ---- additional reference helpers (``runtime/zero/utils.py``) --------------------------------------------------------
class ZeRORuntimeException(Exception):
    pass


def _zero_supported():
    try:
        return _supported()
    except Exception:
        return []


ZERO_SUPPORTED_OPTIMIZERS = _zero_supported()


def is_builtin_type(obj):
    return obj.__class__.__module__ in ("__builtin__", "builtins")


def isinstance_namedtuple(obj) -> bool:
    return isinstance(obj, tuple) and hasattr(obj, "_asdict") and hasattr(obj, "_fields")


def apply_to_tensors_only(function, value, warning_msg_fn=None):
    """Apply ``function`` to every tensor inside nested lists / tuples / namedtuples / dicts, keep everything else."""
    import torch
    if isinstance_namedtuple(value):
        return value.__class__(*(apply_to_tensors_only(function, v, warning_msg_fn) for v in value))
    if isinstance(value, (tuple, list)):
        return value.__class__(apply_to_tensors_only(function, v, warning_msg_fn) for v in value)
    if isinstance(value, dict):
        return {k: apply_to_tensors_only(function, v, warning_msg_fn) for k, v in value.items()}
    if isinstance(value, torch.Tensor):
        return function(value)
    if warning_msg_fn is not None and not is_builtin_type(value):
        from deepspeed_b200.utils import logger
        logger.warning(warning_msg_fn(value))
    return value


def get_mapping_to_flat_buffer(tensors):
    """[(tensor, offset, numel)] of ``tensors`` laid out back to back in one flat buffer."""
    out, offset = [], 0
    for t in tensors:
        out.append((t, offset, t.numel()))
        offset += t.numel()
    return out


def is_zero_param(parameter):
    """Is ``parameter`` managed (sharded) by ZeRO-3?"""
    import torch
    return torch.is_tensor(parameter) and hasattr(parameter, "ds_id")
