"""``CPUAdamBuilder`` (reference ``op_builder/cpu_adam.py``): the op lives in one of the two in-tree native libraries; see
``op_builder/__init__.py``."""
from . import CPUAdamBuilder  # noqa: F401
