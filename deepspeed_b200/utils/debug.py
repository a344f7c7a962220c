"""Debug helpers (reference ``utils/debug.py``): parameter/module name registries and rank-tagged printing."""
import fcntl
import os

module_names = {}
param_names = {}


def debug_clear_module_and_param_names():
    global module_names, param_names
    module_names, param_names = {}, {}


def debug_extract_module_and_param_names(model):
    global module_names, param_names
    module_names = {module: name for name, module in model.named_modules()}
    param_names = {param: name for name, param in model.named_parameters()}


def debug_module2name(module):
    return module_names.get(module, "unknown")


def debug_module2name_id(module):
    return f"name={debug_module2name(module)} id={getattr(module, 'id', id(module))}"


def debug_module2name_class(module):
    return f"name={debug_module2name(module)} {module.__class__.__name__}"


def debug_param2name(param):
    return param_names.get(param, "unknown")


def debug_param2name_id(param):
    return f"name={debug_param2name(param)} id={getattr(param, 'ds_id', id(param))}"


def debug_param2name_id_shape(param):
    return f"name={debug_param2name(param)} id={getattr(param, 'ds_id', id(param))} shape={tuple(param.data.shape)}"


def debug_param2name_id_shape_device(param):
    return f"{debug_param2name_id_shape(param)} device={param.device}"


def debug_param2name_id_numel(param):
    return f"name={debug_param2name(param)} id={getattr(param, 'ds_id', id(param))} numel={param.numel()}"


def debug_param2name_id_shape_status(param):
    return f"{debug_param2name_id_shape(param)} status={getattr(param, 'ds_status', 'n/a')}"


def printflock(*msgs):
    """Print without interleaving across processes (file lock on this source file)."""
    with open(__file__, "r") as fh:
        fcntl.flock(fh, fcntl.LOCK_EX)
        try:
            print(*msgs)
        finally:
            fcntl.flock(fh, fcntl.LOCK_UN)


_fh = None


def log_rank_file(rank, *msgs):
    """Append to ``log_rank_<rank>.txt`` (per-rank trace files for hang debugging)."""
    global _fh
    if _fh is None:
        _fh = open(f"log_rank_{rank}.txt", "w")
    for m in msgs:
        _fh.write(f"{m}\n")
    _fh.flush()


def print_backward_tensors(tensor):

    def walk(fn):
        print(f"Backward tensors in {fn}")
        for nxt in fn.next_functions:
            if nxt[0]:
                t = getattr(nxt[0], "variable", None)
                if t is not None:
                    print(nxt[0], f"Tensor - id: {id(t)}, shape: {t.shape}, data: {t}, grad: {t.grad}")
                walk(nxt[0])

    if hasattr(tensor, "grad_fn") and tensor.grad_fn is not None:
        walk(tensor.grad_fn)
