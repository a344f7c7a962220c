from .policy import Llama2Policy  # noqa: F401
from .model import Llama2InferenceModel  # noqa: F401
from .container import Llama2NonTransformerContainer, Llama2TransformerContainer  # noqa: F401
