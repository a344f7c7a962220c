from .config import DeepSpeedInferenceConfig  # noqa: F401
from .engine import InferenceEngine  # noqa: F401
from .v2 import InferenceEngineV2, RaggedInferenceEngineConfig, build_hf_engine  # noqa: F401,E402
from .v2.config_v2 import DeepSpeedTPConfig  # noqa: F401,E402
from .v2.engine_factory import build_engine_from_ds_checkpoint  # noqa: F401,E402
