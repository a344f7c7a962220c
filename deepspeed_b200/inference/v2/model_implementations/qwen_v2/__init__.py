from .policy import Qwen2Policy  # noqa: F401
from .model import Qwen2InferenceModel  # noqa: F401
from .container import Qwen2NonTransformerContainer, Qwen2TransformerContainer  # noqa: F401
