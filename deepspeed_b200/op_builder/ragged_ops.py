"""``RaggedOpsBuilder`` (reference ``op_builder/ragged_ops.py``): the op lives in one of the two in-tree native libraries; see
``op_builder/__init__.py``."""
from . import RaggedOpsBuilder  # noqa: F401
