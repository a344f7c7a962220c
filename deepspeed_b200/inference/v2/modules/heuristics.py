"""Pick an implementation for each module kind given its config and the engine config (reference
``inference/v2/modules/heuristics.py``)."""
from ..inference_utils import NormTypeEnum
from . import implementations  # noqa: F401  (registers the implementations)
from .configs import (DSEmbeddingsConfig, DSLinearConfig, DSMoEConfig, DSNormConfig, DSSelfAttentionConfig, DSUnembedConfig)
from .interfaces import (DSEmbeddingRegistry, DSLinearRegistry, DSMoERegistry, DSPostNormRegistry, DSPreNormRegistry,
                         DSSelfAttentionRegistry, DSUnembedRegistry)
from .module_registry import ConfigBundle


def _impl_cfg(engine_config, **extra):
    out = dict(extra)
    sm = getattr(engine_config, "state_manager", None)
    if sm is not None and getattr(sm, "max_context", None):
        out.setdefault("max_positions", int(sm.max_context))
    return out


def instantiate_attention(attention_config: DSSelfAttentionConfig, engine_config=None):
    return DSSelfAttentionRegistry.instantiate_config(
        ConfigBundle(name="dense_blocked_attention", config=attention_config, implementation_config=_impl_cfg(engine_config)))


def instantiate_embed(embed_config: DSEmbeddingsConfig, engine_config=None):
    return DSEmbeddingRegistry.instantiate_config(ConfigBundle(name="ragged_embedding", config=embed_config))


def instantiate_linear(linear_config: DSLinearConfig, engine_config=None):
    """Quantised implementation when the engine (or the module config) asks for a weight-only mode, else the BLAS one."""
    mode = linear_config.quantization_mode or getattr(getattr(engine_config, "quantization", None), "quantization_mode", None)
    if mode is None:
        return DSLinearRegistry.instantiate_config(ConfigBundle(name="blas_fp_linear", config=linear_config))
    cfg = linear_config.model_copy(update={"quantization_mode": mode})
    return DSLinearRegistry.instantiate_config(ConfigBundle(name="quantized_wf6af16_linear", config=cfg))


def instantiate_moe(moe_config: DSMoEConfig, engine_config=None):
    return DSMoERegistry.instantiate_config(ConfigBundle(name="cutlass_multi_gemm_moe", config=moe_config))


def instantiate_post_norm(norm_config: DSNormConfig, engine_config=None):
    return DSPostNormRegistry.instantiate_config(ConfigBundle(name="cuda_post_ln", config=norm_config))


def instantiate_pre_norm(norm_config: DSNormConfig, engine_config=None):
    name = "cuda_pre_rms" if NormTypeEnum(norm_config.type) == NormTypeEnum.RMSNorm else "cuda_pre_ln"
    return DSPreNormRegistry.instantiate_config(ConfigBundle(name=name, config=norm_config))


def instantiate_unembed(unembed_config: DSUnembedConfig, engine_config=None):
    return DSUnembedRegistry.instantiate_config(ConfigBundle(name="ragged_unembed", config=unembed_config))
