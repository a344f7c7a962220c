from .policy import Phi3Policy  # noqa: F401
from .model import Phi3InferenceModel  # noqa: F401
from .container import Phi3NonTransformerContainer, Phi3TransformerContainer  # noqa: F401
