from .config import DeepSpeedInferenceConfig  # noqa: F401
from .ds_transformer import DeepSpeedTransformerInference  # noqa: F401
from .moe_inference import DeepSpeedMoEInference, DeepSpeedMoEInferenceConfig  # noqa: F401
