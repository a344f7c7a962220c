"""``CPUAdagradBuilder`` (reference ``op_builder/cpu_adagrad.py``): the op lives in one of the two in-tree native libraries; see
``op_builder/__init__.py``."""
from . import CPUAdagradBuilder  # noqa: F401
