"""Sequence / context parallelism: DeepSpeed-Ulysses all-to-all, FPDT chunked attention, ring attention."""
from .layer import DistributedAttention, single_all_to_all  # noqa: F401
