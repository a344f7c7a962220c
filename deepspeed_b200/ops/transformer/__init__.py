from .transformer import DeepSpeedTransformerLayer, DeepSpeedTransformerConfig, TransformerConfig  # noqa: F401
from .inference import DeepSpeedInferenceConfig, DeepSpeedTransformerInference  # noqa: F401
from .inference.moe_inference import DeepSpeedMoEInference, DeepSpeedMoEInferenceConfig  # noqa: F401,E402
