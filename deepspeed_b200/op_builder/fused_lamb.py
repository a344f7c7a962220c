"""``FusedLambBuilder`` (reference ``op_builder/fused_lamb.py``): the op lives in one of the two in-tree native libraries; see
``op_builder/__init__.py``."""
from . import FusedLambBuilder  # noqa: F401
