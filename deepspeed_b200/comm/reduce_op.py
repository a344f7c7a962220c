"""Reduce-op enum mirrored from torch.distributed (reference: ``comm/reduce_op.py``)."""
from enum import Enum


class ReduceOp(Enum):
    SUM = 0
    PRODUCT = 1
    MIN = 2
    MAX = 3
    BAND = 4
    BOR = 5
    BXOR = 6
    AVG = 7
    UNUSED = 8


def to_torch(op):
    import torch.distributed as dist
    if isinstance(op, ReduceOp):
        return {
            ReduceOp.SUM: dist.ReduceOp.SUM,
            ReduceOp.PRODUCT: dist.ReduceOp.PRODUCT,
            ReduceOp.MIN: dist.ReduceOp.MIN,
            ReduceOp.MAX: dist.ReduceOp.MAX,
            ReduceOp.BAND: dist.ReduceOp.BAND,
            ReduceOp.BOR: dist.ReduceOp.BOR,
            ReduceOp.BXOR: dist.ReduceOp.BXOR,
            ReduceOp.AVG: dist.ReduceOp.AVG,
        }[op]
    return op
