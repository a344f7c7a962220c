from .policy import FalconPolicy  # noqa: F401
from .model import FalconInferenceModel  # noqa: F401
from .container import FalconNewArchTransformerContainer, FalconNonTransformerContainer, FalconTransformerContainer  # noqa: F401
