from .compressed import CompressedBackend, NcclBackend, MpiBackend, HcclBackend  # noqa: F401
from .coalesced_collectives import reduce_scatter_coalesced, all_to_all_quant_reduce, all_to_all_loco_quant_reduce  # noqa: F401
