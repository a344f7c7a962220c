from .policy import QwenPolicy  # noqa: F401
from .model import QwenInferenceModel  # noqa: F401
from .container import QwenNonTransformerContainer, QwenTransformerContainer  # noqa: F401
