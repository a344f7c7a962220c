"""``QuantizerBuilder`` (reference ``op_builder/quantizer.py``): the op lives in one of the two in-tree native libraries; see
``op_builder/__init__.py``."""
from . import QuantizerBuilder  # noqa: F401
