"""Communication layer: torch.distributed facade + symmetric-memory (NVLink peer) layer."""
from .reduce_op import ReduceOp  # noqa: F401
from .comm import *  # noqa: F401,F403
from .comm import (init_distributed, is_initialized, get_rank, get_world_size, get_local_rank, barrier,  # noqa: F401
                   comms_logger, configure, log_summary)
