"""``FusedAdamBuilder`` (reference ``op_builder/fused_adam.py``): the op lives in one of the two in-tree native libraries; see
``op_builder/__init__.py``."""
from . import FusedAdamBuilder  # noqa: F401
