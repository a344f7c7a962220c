"""Accessors that tolerate the different spellings model-parallel units (Megatron old/new, our own grid) use for the
same group (reference ``utils/bwc.py``).  ``mpu=None`` falls back to the framework's own grid in ``utils.groups``."""


def _first(mpu, names, default):
    for n in names:
        fn = getattr(mpu, n, None)
        if fn is not None:
            return fn()
    return default()


def bwc_tensor_model_parallel_rank(mpu=None):
    from . import groups
    if mpu is None:
        return groups._get_model_parallel_rank()
    return _first(mpu, ("get_tensor_model_parallel_rank", "get_slice_parallel_rank", "get_model_parallel_rank"), lambda: 0)


def bwc_tensor_model_parallel_world_size(mpu=None):
    from . import groups
    if mpu is None:
        return groups._get_model_parallel_world_size()
    return _first(mpu, ("get_tensor_model_parallel_world_size", "get_slice_parallel_world_size",
                        "get_model_parallel_world_size"), lambda: 1)


def bwc_tensor_model_parallel_group(mpu=None):
    from . import groups
    if mpu is None:
        return groups._get_model_parallel_group()
    return _first(mpu, ("get_tensor_model_parallel_group", "get_slice_parallel_group", "get_model_parallel_group"),
                  lambda: None)


def bwc_pipeline_parallel_world_size(mpu=None):
    from . import groups
    if mpu is None:
        return groups._get_pipe_parallel_world_size()
    return _first(mpu, ("get_pipeline_model_parallel_world_size", "get_pipe_parallel_world_size"), lambda: 1)


def bwc_pipeline_parallel_group(mpu=None):
    from . import groups
    if mpu is None:
        return groups._get_pipe_parallel_group()
    return _first(mpu, ("get_pipeline_model_parallel_group", "get_pipe_parallel_group"), lambda: None)
