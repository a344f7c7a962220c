

from . import adagrad, adam, fp_quantizer, lamb, lion, sparse_attention  # noqa: F401,E402
from .transformer import DeepSpeedTransformerConfig, DeepSpeedTransformerLayer  # noqa: F401,E402
from ..git_version_info import compatible_ops as __compatible_ops__  # noqa: F401,E402  (reference ``ops/__init__.py:15``)
