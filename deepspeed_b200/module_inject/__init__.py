from .layers import LinearAllreduce, LinearLayer, LmHeadLinearAllreduce, RowParallel, ColumnParallel  # noqa: F401
from .auto_tp import AutoTP, ReplaceWithTensorSlicing, tp_model_init  # noqa: F401
from .replace_module import replace_transformer_layer, generic_injection  # noqa: F401
