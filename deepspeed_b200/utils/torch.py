"""torch version gates (reference ``utils/torch.py``)."""
import torch
from packaging import version as _v


def required_torch_version(min_version=None, max_version=None):
    assert min_version or max_version, "Must provide a min_version or max_version argument"
    cur = _v.parse(torch.__version__.split("+")[0])
    if min_version and _v.parse(str(min_version)) > cur:
        return False
    if max_version and _v.parse(str(max_version)) < cur:
        return False
    return True


def register_grad_hook(param, hook):
    """Run ``hook(param)`` once the gradient of ``param`` has been accumulated."""
    return param.register_post_accumulate_grad_hook(hook)


def jit_script_compat(fn):
    return fn
