from .adam import OnebitAdam  # noqa: F401
from .zoadam import ZeroOneAdam  # noqa: F401
from .lamb import OnebitLamb  # noqa: F401


def build_onebit_optimizer(name, model_parameters, params, engine):
    name = name.lower()
    params = dict(params or {})
    params.pop("comm_backend_name", None)
    cls = {"onebitadam": OnebitAdam, "zerooneadam": ZeroOneAdam, "onebitlamb": OnebitLamb}[name]
    return cls(model_parameters, deepspeed=engine, **params)
