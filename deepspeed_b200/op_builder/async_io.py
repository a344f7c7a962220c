"""``AsyncIOBuilder`` (reference ``op_builder/async_io.py``): the op lives in one of the two in-tree native libraries; see
``op_builder/__init__.py``."""
from . import AsyncIOBuilder  # noqa: F401
