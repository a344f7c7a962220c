"""Launcher-environment discovery and message-size helpers for the comms logger (reference ``comm/utils.py``)."""
import inspect
import os


def _env_int(names, what):
    for n in names:
        v = os.environ.get(n)
        if v is not None:
            return int(v)
    raise RuntimeError(f"{what} could not be determined: none of {', '.join(names)} is set (launch with the deepspeed "
                       "launcher, torchrun or mpirun)")


def get_local_rank_from_launcher():
    return _env_int(("LOCAL_RANK", "OMPI_COMM_WORLD_LOCAL_RANK", "MV2_COMM_WORLD_LOCAL_RANK", "SLURM_LOCALID"), "local rank")


def get_world_rank_from_launcher():
    return _env_int(("RANK", "OMPI_COMM_WORLD_RANK", "PMI_RANK", "SLURM_PROCID"), "world rank")


def get_world_size_from_launcher():
    return _env_int(("WORLD_SIZE", "OMPI_COMM_WORLD_SIZE", "PMI_SIZE", "SLURM_NTASKS"), "world size")


def get_default_args(func):
    return {k: v.default for k, v in inspect.signature(func).parameters.items() if v.default is not inspect.Parameter.empty}


_TENSOR_ARG_NAMES = ("tensor", "tensors", "input_list", "input_tensor_list", "input_tensor", "input", "output_tensor")


def get_tensor_position(func):
    """Index of the positional argument that carries the payload tensor (or -1)."""
    params = list(inspect.signature(func).parameters)
    for name in _TENSOR_ARG_NAMES:
        if name in params:
            return params.index(name)
    return -1


def get_tensor_kwarg(func, kwargs):
    merged = {**get_default_args(func), **kwargs}
    for name in _TENSOR_ARG_NAMES:
        if merged.get(name) is not None:
            return merged[name]
    return None


def get_msg_size_from_args(func, *args, **kwargs):
    """Payload bytes of a collective call, wherever the tensor (or tensor list) was passed."""
    pos = get_tensor_position(func)
    arg = args[pos] if 0 <= pos < len(args) else get_tensor_kwarg(func, kwargs)
    if arg is None:
        return 0
    if isinstance(arg, (list, tuple)):
        return sum(t.element_size() * t.nelement() for t in arg)
    return arg.element_size() * arg.nelement()


def get_debug_log_name(func_args, debug):
    return func_args["log_name"] + " | [Caller Func: " + inspect.stack()[2][3] + "]" if debug else func_args["log_name"]
