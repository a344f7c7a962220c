"""``SpatialInferenceBuilder`` (reference ``op_builder/spatial_inference.py``): the op lives in one of the two in-tree native libraries; see
``op_builder/__init__.py``."""
from . import SpatialInferenceBuilder  # noqa: F401
