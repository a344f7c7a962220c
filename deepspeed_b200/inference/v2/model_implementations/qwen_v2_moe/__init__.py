from .policy import Qwen2MoePolicy  # noqa: F401
from .model import Qwen2MoeInferenceModel  # noqa: F401
from .container import Qwen2MoeNonTransformerContainer, Qwen2MoeTransformerContainer  # noqa: F401
