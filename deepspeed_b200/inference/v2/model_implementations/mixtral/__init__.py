from .policy import MixtralPolicy  # noqa: F401
from .model import MixtralInferenceModel  # noqa: F401
from .container import MixtralNonTransformerContainer, MixtralTransformerContainer  # noqa: F401
