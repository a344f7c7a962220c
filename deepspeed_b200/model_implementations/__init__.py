"""Inference model wrappers (reference ``model_implementations/``): CUDA-graphed diffusers / CLIP wrappers and the
per-family fused transformer layers used by kernel injection."""
from .diffusers.unet import DSUNet  # noqa: F401
from .diffusers.vae import DSVAE  # noqa: F401
from .transformers.clip_encoder import DSClipEncoder  # noqa: F401
from .transformers.ds_transformer import DeepSpeedTransformerInference  # noqa: F401
from .transformers.ds_bert import DeepSpeedBERTInference  # noqa: F401
from .transformers.ds_bloom import DeepSpeedBloomInference  # noqa: F401
from .transformers.ds_gpt import DeepSpeedGPTInference  # noqa: F401
from .transformers.ds_llama2 import DeepSpeedLlama2Inference  # noqa: F401
from .transformers.ds_megatron_gpt import DeepSpeedMegatronGPTInference  # noqa: F401
from .transformers.ds_opt import DeepSpeedOPTInference  # noqa: F401
