"""Channel-last bias-add fusions for diffusion UNet / VAE (reference ``csrc/spatial`` N13, ``SpatialInferenceBuilder``)."""
from deepspeed_b200.ops.kernels.misc_ops import nhwc_bias_add as _k


def nhwc_bias_add(activation, bias, other=None, other_bias=None):
    """``activation`` is a channels-last tensor ([N,H,W,C] memory order)."""
    if activation.dim() == 4 and activation.is_contiguous(memory_format=__import__("torch").channels_last):
        x = activation.permute(0, 2, 3, 1)
        o = other.permute(0, 2, 3, 1) if other is not None else None
        return _k(x, bias, o, other_bias).permute(0, 3, 1, 2)
    return _k(activation, bias, other, other_bias)


def nhwc_bias_add_add(activation, bias, other):
    return nhwc_bias_add(activation, bias, other)


def nhwc_bias_add_bias_add(activation, bias, other, other_bias):
    return nhwc_bias_add(activation, bias, other, other_bias)
