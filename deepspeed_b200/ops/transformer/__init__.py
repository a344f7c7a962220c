from .transformer import DeepSpeedTransformerLayer, DeepSpeedTransformerConfig, TransformerConfig  # noqa: F401
from .inference import DeepSpeedInferenceConfig, DeepSpeedTransformerInference  # noqa: F401
