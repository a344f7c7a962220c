from .logging import logger, log_dist, print_json_dist, warning_once, set_log_level  # noqa: F401
from . import groups  # noqa: F401
from .timer import SynchronizedWallClockTimer, ThroughputTimer, NoopTimer  # noqa: F401
from .nvtx import instrument_w_nvtx  # noqa: F401
from .init_on_device import OnDevice  # noqa: F401
from .tensor_fragment import (safe_get_full_fp32_param, safe_get_full_grad, safe_get_full_optimizer_state,  # noqa: F401
                              safe_set_full_fp32_param, safe_set_full_optimizer_state, safe_get_local_fp32_param,
                              safe_get_local_grad, safe_get_local_optimizer_state, safe_set_local_fp32_param,
                              safe_set_local_optimizer_state, safe_set_full_grad, safe_set_local_grad)
from .z3_leaf_module import set_z3_leaf_modules, unset_z3_leaf_modules, get_z3_leaf_modules, z3_leaf_module  # noqa: F401
