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
