"""``torch.compile`` interplay (reference ``runtime/compiler.py``).  This framework never uses a tracing compiler on its
hot path; these helpers exist so framework internals can opt out of a user's ``torch.compile`` region and user code can
query compile state."""
import torch

try:
    from torch.compiler import is_compiling as _torch_is_compiling
except ImportError:  # very old torch
    _torch_is_compiling = lambda: False  # noqa: E731


def is_compile_supported():
    return hasattr(torch, "compiler") and hasattr(torch.compiler, "disable")


def disable(func):
    return torch.compiler.disable(func) if is_compile_supported() else func


def is_compiling():
    return _torch_is_compiling()
