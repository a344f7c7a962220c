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

    # ---- constructor wrappers (reference ``init_on_device.py:42-62``) -------------------------------------------------
    # The context itself does not need them (``torch.device`` + default dtype cover every factory function); they are
    # offered for code that wraps individual constructors explicitly.
    def fp_tensor_constructor(self, fn, target_fp_dtype):
        """Wrap a factory (``torch.empty`` ...) so results land on ``self.device`` and float results use ``target_fp_dtype``."""

        def wrapped_fn(*args, **kwargs):
            if kwargs.get("device") is None:
                kwargs["device"] = self.device
            t = fn(*args, **kwargs)
            return t.to(target_fp_dtype) if t.is_floating_point() else t

        return wrapped_fn

    def get_new_tensor_fn_for_dtype(self, dtype):
        """Replacement for ``Tensor.__new__``-style construction: ``new_tensor(cls, *sizes)``."""

        def new_tensor(cls, *args):
            t = torch.empty(0, device=self.device).new_empty(*args)
            return t.to(dtype) if t.is_floating_point() else t

        return new_tensor
