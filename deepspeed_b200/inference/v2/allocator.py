"""View allocation helpers (reference ``inference/v2/allocator.py``)."""
from functools import reduce
from typing import Iterable

import torch

from deepspeed_b200.accelerator import get_accelerator


class Allocator:
    """Reuse one pre-allocated buffer for differently-shaped activations: ``empty_from(buf, shape)`` is a view of its head."""
    cache = {}

    @staticmethod
    def empty_from(tensor: torch.Tensor, shape: Iterable[int]) -> torch.Tensor:
        shape = tuple(shape)
        n = reduce(lambda a, b: a * b, shape, 1)
        if n == 0:
            raise ValueError("Cannot create empty tensor with size 0")
        if n > tensor.numel():
            raise ValueError(f"buffer of {tensor.numel()} elements cannot hold shape {shape}")
        return tensor.flatten()[:n].view(shape)


empty_from = Allocator.empty_from


def on_device(method):
    """Decorator for parameter-transform methods: the returned tensor lands on the current accelerator."""

    def wrapped(self, *args, **kwargs):
        out = method(self, *args, **kwargs)
        dev = get_accelerator().current_device_name() if get_accelerator().is_available() else "cpu"
        return out.to(dev) if isinstance(out, torch.Tensor) else out

    return wrapped
