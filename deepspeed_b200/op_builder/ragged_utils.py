"""``RaggedUtilsBuilder`` (reference ``op_builder/ragged_utils.py``): the op lives in one of the two in-tree native libraries; see
``op_builder/__init__.py``."""
from . import RaggedUtilsBuilder  # noqa: F401
