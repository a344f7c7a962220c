"""``StochasticTransformerBuilder`` (reference ``op_builder/stochastic_transformer.py``): the op lives in one of the two in-tree native libraries; see
``op_builder/__init__.py``."""
from . import StochasticTransformerBuilder  # noqa: F401
