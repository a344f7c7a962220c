"""``TransformerBuilder`` (reference ``op_builder/transformer.py``): the op lives in one of the two in-tree native libraries; see
``op_builder/__init__.py``."""
from . import TransformerBuilder  # noqa: F401
