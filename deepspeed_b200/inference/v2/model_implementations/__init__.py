from .arch import ArchSpec, arch_from_hf_config, SUPPORTED_MODEL_TYPES  # noqa: F401
from .ragged_transformer import RaggedTransformer  # noqa: F401
from .weights import load_hf_weights, weights_from_b200_model  # noqa: F401
