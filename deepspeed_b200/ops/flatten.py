"""``flatten`` / ``unflatten`` (reference ``csrc/utils/flatten_unflatten.cpp`` N15, ``UtilsBuilder``)."""
from typing import List

import torch


def flatten(tensors: List[torch.Tensor]) -> torch.Tensor:
    if not tensors:
        return torch.empty(0)
    out = torch.empty(sum(t.numel() for t in tensors), dtype=tensors[0].dtype, device=tensors[0].device)
    torch._foreach_copy_(list(unflatten(out, tensors)), [t.detach() for t in tensors]) if hasattr(torch, "_foreach_copy_") \
        else [v.copy_(t) for v, t in zip(unflatten(out, tensors), tensors)]
    return out


def unflatten(flat: torch.Tensor, tensors: List[torch.Tensor]):
    views, off = [], 0
    for t in tensors:
        n = t.numel()
        views.append(flat.narrow(0, off, n).view(t.shape))
        off += n
    return tuple(views)


class UtilsBuilder:

    def load(self, verbose=False):
        import sys
        return sys.modules[__name__]
