from .policy import PhiPolicy  # noqa: F401
from .model import PhiInferenceModel  # noqa: F401
from .container import PhiNonTransformerContainer, PhiTransformerContainer  # noqa: F401
