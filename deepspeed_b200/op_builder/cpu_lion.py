"""``CPULionBuilder`` (reference ``op_builder/cpu_lion.py``): the op lives in one of the two in-tree native libraries; see
``op_builder/__init__.py``."""
from . import CPULionBuilder  # noqa: F401
