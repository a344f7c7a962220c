"""``RandomLTDBuilder`` (reference ``op_builder/random_ltd.py``): the op lives in one of the two in-tree native libraries; see
``op_builder/__init__.py``."""
from . import RandomLTDBuilder  # noqa: F401
