"""NVTX range decorator (reference: ``utils/nvtx.py:11 instrument_w_nvtx``)."""
import functools
import os

_enabled = os.environ.get("DSB200_NVTX", "0") == "1"


def enable_nvtx(flag=True):
    global _enabled
    _enabled = flag


def instrument_w_nvtx(func):
    """Push/pop an NVTX range named after ``func`` when ``DSB200_NVTX=1`` (zero cost otherwise)."""

    @functools.wraps(func)
    def wrapped(*args, **kwargs):
        if not _enabled:
            return func(*args, **kwargs)
        from deepspeed_b200.accelerator import get_accelerator
        acc = get_accelerator()
        acc.range_push(func.__qualname__)
        try:
            return func(*args, **kwargs)
        finally:
            acc.range_pop()

    return wrapped
