"""``InferenceCutlassBuilder`` (reference ``op_builder/inference_cutlass_builder.py``): the op lives in one of the two in-tree native libraries; see
``op_builder/__init__.py``."""
from . import InferenceCutlassBuilder  # noqa: F401
