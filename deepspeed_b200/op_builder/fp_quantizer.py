"""``FPQuantizerBuilder`` (reference ``op_builder/fp_quantizer.py``): the op lives in one of the two in-tree native libraries; see
``op_builder/__init__.py``."""
from . import FPQuantizerBuilder  # noqa: F401
