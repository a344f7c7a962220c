"""Kernel-object protocol of the ragged engine (reference ``inference/v2/kernels/ds_kernel.py``): construction validates
the configuration (dtype, shapes) once, ``__call__`` launches without allocating."""
from abc import ABC, abstractmethod

import torch

SUPPORTED_DTYPES = (torch.float16, torch.bfloat16)


class DSKernelBase(ABC):

    @abstractmethod
    def __init__(self, *args, **kwargs):
        raise NotImplementedError()

    @abstractmethod
    def __call__(self, *args, **kwargs):
        raise NotImplementedError()


def check_dtype(dtype, what, allow_fp32=True):
    ok = SUPPORTED_DTYPES + ((torch.float32, ) if allow_fp32 else ())
    if dtype not in ok:
        raise ValueError(f"Unsupported data type {dtype} for {what}; supported: {ok}")
