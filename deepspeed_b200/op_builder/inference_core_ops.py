"""``InferenceCoreBuilder`` (reference ``op_builder/inference_core_ops.py``): the op lives in one of the two in-tree native libraries; see
``op_builder/__init__.py``."""
from . import InferenceCoreBuilder  # noqa: F401
