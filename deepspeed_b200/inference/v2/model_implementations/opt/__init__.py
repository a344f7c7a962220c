from .policy import OPTPolicy  # noqa: F401
from .model import OPTInferenceModel  # noqa: F401
from .container import OPTNonTransformerContainer, OPTTransformerContainer  # noqa: F401
