"""``EvoformerAttnBuilder`` (reference ``op_builder/evoformer_attn.py``): the op lives in one of the two in-tree native libraries; see
``op_builder/__init__.py``."""
from . import EvoformerAttnBuilder  # noqa: F401
