"""``GDSBuilder`` (reference ``op_builder/gds.py``): the op lives in one of the two in-tree native libraries; see
``op_builder/__init__.py``."""
from . import GDSBuilder  # noqa: F401
