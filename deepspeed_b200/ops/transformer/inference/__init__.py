from .config import DeepSpeedInferenceConfig  # noqa: F401
from .ds_transformer import DeepSpeedTransformerInference  # noqa: F401
from .moe_inference import DeepSpeedMoEInference, DeepSpeedMoEInferenceConfig  # noqa: F401


# ---- names of the reference's v1 inference op bindings (ops/transformer/inference/op_binding/*) ---------------------
def pre_rms_norm(x, residual, gamma, eps=1e-6):
    """``residual += x; return rmsnorm(residual), residual`` (reference ``pre_rms_norm``)."""
    from deepspeed_b200.ops.kernels.transformer_ops import rms_norm
    return rms_norm(x, gamma, eps, residual=residual)


def attn_softmax_v2(scores, mask=None, alibi=None, scale=1.0, causal=True, window=0):
    from deepspeed_b200.ops.kernels.misc_ops import attn_softmax
    return attn_softmax(scores, mask=mask, alibi=alibi, scale=scale, causal=causal, window=window)


def ds_softmax_context(q, k, v, scale=None, causal=True):
    """Fused scores -> masked softmax -> context (reference ``softmax_context``): flash SDPA on device."""
    import torch.nn.functional as F
    return F.scaled_dot_product_attention(q, k, v, is_causal=causal, scale=scale)


def moe_res_matmul(moe_out, mlp_out, coef):
    """Residual-MoE mixing: ``mlp_out * coef[...,0] + moe_out * coef[...,1]`` (reference ``moe_res_matmul``)."""
    return mlp_out * coef[..., 0:1] + moe_out * coef[..., 1:2]
