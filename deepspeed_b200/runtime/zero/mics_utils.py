"""MiCS communicator construction (reference ``runtime/zero/mics_utils.py``); lives with the rest of MiCS in ``mics.py``."""
from .mics import (MiCS_CommGroups, _generate_mics_config, create_mics_comm_groups, hierarchical_all_gather,  # noqa: F401
                   hierarchy_layout, mics_rank_layout, scale_tensors)
