"""``OnDevice`` -- construct a model on ``meta`` or a real device with a chosen dtype.

Parity target: reference ``utils/init_on_device.py``.  Implemented with ``torch.device`` as a
context manager plus a default-dtype override instead of patching tensor constructors.
"""
import contextlib

import torch


class OnDevice(contextlib.AbstractContextManager):

    def __init__(self, dtype=None, device="meta", enabled=True):
        self.dtype = dtype
        self.device = device
        self.enabled = enabled
        self._stack = None

    def __enter__(self):
        if not self.enabled:
            return self
        self._stack = contextlib.ExitStack()
        self._stack.enter_context(torch.device(self.device))
        if self.dtype is not None:
            prev = torch.get_default_dtype()
            torch.set_default_dtype(self.dtype)
            self._stack.callback(torch.set_default_dtype, prev)
        return self

    def __exit__(self, *exc):
        if self._stack is not None:
            self._stack.close()
            self._stack = None
        return False
