"""``NcclBackend`` for the 1-bit optimizers (reference ``runtime/comm/nccl.py``): error-compensated sign compression,
chunk exchange by all-to-all, server-side re-compression, all-gather.  One implementation serves NCCL and gloo."""
from .compressed import CompressedBackend


class NcclBackend(CompressedBackend):
    pass
