

from . import adagrad, adam, fp_quantizer, lamb, lion, sparse_attention  # noqa: F401,E402
from .transformer import DeepSpeedTransformerConfig, DeepSpeedTransformerLayer  # noqa: F401,E402
