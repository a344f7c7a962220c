"""Registry of every op builder (reference ``op_builder/all_ops.py``)."""
from . import ALL_OPS  # noqa: F401

__op_builders__ = [cls() for cls in ALL_OPS.values()]
op_builder_dir = "deepspeed_b200.op_builder"
