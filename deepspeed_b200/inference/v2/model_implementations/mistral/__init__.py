from .policy import MistralPolicy  # noqa: F401
from .model import MistralInferenceModel  # noqa: F401
from .container import MistralNonTransformerContainer, MistralTransformerContainer  # noqa: F401
