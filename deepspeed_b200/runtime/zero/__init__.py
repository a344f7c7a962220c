"""ZeRO: sharded optimizer states / gradients / parameters (stages 1-3), offload tiers, ZeRO++."""
from .config import DeepSpeedZeroConfig, ZeroStageEnum, get_zero_config  # noqa: F401
from .offload_config import (DeepSpeedZeroOffloadOptimizerConfig, DeepSpeedZeroOffloadParamConfig,  # noqa: F401
                             OffloadDeviceEnum, OffloadStateTypeEnum)


def __getattr__(name):
    # Heavy modules are imported lazily so `import deepspeed_b200` stays fast.
    if name in ("Init", "GatheredParameters", "register_external_parameter", "unregister_external_parameter",
                "ZeroParamStatus", "ZeroParamType", "shutdown_init_context", "restore_init_context"):
        from . import partition_parameters as pp
        return getattr(pp, name)
    if name in ("TiledLinear", "TiledLinearReturnBias"):
        from . import tiling
        return getattr(tiling, name)
    if name in ("MiCS_Init", "MiCS_Optimizer"):
        from . import mics
        return getattr(mics, name)
    if name in ("estimate_zero2_model_states_mem_needs_all_live", "estimate_zero2_model_states_mem_needs_all_cold",
                "estimate_zero3_model_states_mem_needs_all_live", "estimate_zero3_model_states_mem_needs_all_cold",
                "estimate_zero2_model_states_mem_needs", "estimate_zero3_model_states_mem_needs"):
        from . import mem_estimator
        return getattr(mem_estimator, name)
    raise AttributeError(name)


import contextlib as _contextlib  # noqa: E402


@_contextlib.contextmanager
def unwrap_model_for_generation(model):
    """Gather every ZeRO-3 parameter of ``model`` for the duration of a ``generate()`` call (no per-layer fetches inside
    the autoregressive loop), then re-partition."""
    from .partition_parameters import GatheredParameters, is_zero_param
    params = [p for p in model.parameters() if is_zero_param(p)]
    if not params:
        yield model
        return
    zo = None
    for p in params:
        ref = getattr(p, "_ds_zero", None)
        zo = ref() if ref is not None else None
        if zo is not None:
            break
    if zo is not None and hasattr(zo, "gather_all"):
        zo.gather_all()
        try:
            yield model
        finally:
            zo.release_all()
        return
    with GatheredParameters(params):
        yield model
