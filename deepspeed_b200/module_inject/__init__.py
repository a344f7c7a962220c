from .layers import LinearAllreduce, LinearLayer, LmHeadLinearAllreduce, RowParallel, ColumnParallel  # noqa: F401
from .auto_tp import AutoTP, ReplaceWithTensorSlicing, tp_model_init  # noqa: F401
from .replace_module import replace_transformer_layer, generic_injection  # noqa: F401
from .replace_module import revert_transformer_layer  # noqa: F401,E402
from .module_quantize import quantize_transformer_layer  # noqa: F401,E402
from .policy import DSPolicy  # noqa: F401,E402
from .containers import HFBertLayerPolicy  # noqa: F401,E402
from .layers import EmbeddingLayer, Normalize  # noqa: F401,E402
from .replace_module import GroupQuantizer  # noqa: F401,E402


def set_autotp_mode(training=False):
    from . import layers
    layers.AUTOTP_TRAINING_MODE = bool(training)
