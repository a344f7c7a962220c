"""Per-family injection containers + policies (reference ``module_inject/containers/__init__.py``)."""
from .base import BaseTransformerContainer, InjectedLayer  # noqa: F401
from .base_moe import BaseTransformerMoEContainer  # noqa: F401
from .bert import DS_BERTContainer, HFBertLayerPolicy  # noqa: F401
from .bloom import DS_BloomContainer, BLOOMLayerPolicy  # noqa: F401
from .clip import DS_CLIPContainer, HFCLIPLayerPolicy  # noqa: F401
from .distil_bert import DS_DistilBERTContainer, HFDistilBertLayerPolicy  # noqa: F401
from .gpt2 import DS_GPT2Container, HFGPT2LayerPolicy  # noqa: F401
from .gptj import DS_GPTJContainer, HFGPTJLayerPolicy  # noqa: F401
from .gptneo import DS_GPTNEOContainer, HFGPTNEOLayerPolicy  # noqa: F401
from .gptneox import DS_GPTNEOXContainer, GPTNEOXLayerPolicy  # noqa: F401
from .internlm import DS_InternLMContainer, InternLMLayerPolicy  # noqa: F401
from .llama import DS_LLAMAContainer, LLAMALayerPolicy  # noqa: F401
from .llama2 import DS_LLAMA2Container, LLAMA2LayerPolicy  # noqa: F401
from .megatron_gpt import DS_MegatronGPTContainer, MegatronLayerPolicy  # noqa: F401
from .megatron_gpt_moe import DS_MegatronGPTMoEContainer, MegatronMoELayerPolicy  # noqa: F401
from .opt import DS_OPTContainer, HFOPTLayerPolicy  # noqa: F401
from .unet import UNetPolicy  # noqa: F401
from .vae import VAEPolicy  # noqa: F401
