"""Comms logger (reference ``utils/comms_logging.py``); implemented next to the comm facade in ``comm/comms_logging.py``."""
from deepspeed_b200.comm.comms_logging import *  # noqa: F401,F403
from deepspeed_b200.comm.comms_logging import CommsLogger  # noqa: F401
from deepspeed_b200.comm.comms_logging import calc_bw_log as _calc_bw  # noqa: E402


def get_caller_func(frame=3):
    """Name of the function ``frame`` levels up the stack (who issued the collective)."""
    import sys
    return sys._getframe(frame).f_code.co_name


def print_rank_0(message):
    from deepspeed_b200 import comm as dist
    if not dist.is_initialized() or dist.get_rank() == 0:
        print(message)


def convert_size(size_bytes):
    """``1536 -> "1.5 KB"``."""
    if size_bytes == 0:
        return "0B"
    units = ("B", "KB", "MB", "GB", "TB", "PB", "EB", "ZB", "YB")
    i = min(len(units) - 1, (int(size_bytes).bit_length() - 1) // 10)
    return f"{round(size_bytes / (1 << (10 * i)), 2)} {units[i]}"


def calc_bw_log(comm_op, size, duration):
    """Reference signature (``utils/comms_logging.py:34``): ``duration`` in ms over the default world → ``(algbw, busbw,
    size)`` with the bandwidths in Gbit/s."""
    from deepspeed_b200 import comm as dist
    world = dist.get_world_size() if dist.is_initialized() else 1
    total, alg_gBps, bus_gBps = _calc_bw(comm_op, size, duration, world)
    return alg_gBps * 8, bus_gBps * 8, total
