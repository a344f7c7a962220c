"""``InferenceBuilder`` (reference ``op_builder/transformer_inference.py``): the op lives in one of the two in-tree native libraries; see
``op_builder/__init__.py``."""
from . import InferenceBuilder  # noqa: F401
