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

from .tensor_fragment import (get_hp_fragment_mapping, safe_get_full_fp32_param as get_full_hp_param,  # noqa: F401,E402
                              safe_get_full_grad as get_full_hp_grad, safe_set_full_fp32_param as set_full_hp_param,
                              safe_set_full_grad as set_full_hp_grad)
from .tensor_fragment import fragment_address, map_to_flat_opt_states, tensor_fragment  # noqa: F401,E402
from .mixed_precision_linkage import lazy_init_hp_params_optimizer_state, link_hp_params  # noqa: F401,E402
from .numa import get_numactl_cmd  # noqa: F401,E402
from .z3_leaf_module import set_z3_leaf_module, z3_leaf_parameter  # noqa: F401,E402
from .comms_logging import get_caller_func  # noqa: F401,E402


def __getattr__(name):
    if name == "RepeatingLoader":
        from deepspeed_b200.runtime.dataloader import RepeatingLoader
        return RepeatingLoader
    raise AttributeError(name)
