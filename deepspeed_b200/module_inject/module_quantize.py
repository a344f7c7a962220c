"""In-place int8 quantisation of the linear weights inside transformer layers (reference
``module_inject/module_quantize.py:quantize_transformer_layer``)."""
import torch
from torch import nn


def quantize_transformer_layer(orig_layer_impl, model, megatron=False, preln=False):
    """For every ``orig_layer_impl`` instance, replace each 2-D linear weight by its symmetric per-tensor int8 codes
    (the module keeps ``weight.scale`` so callers can dequantise); returns the model."""

    def quantize_weight(weight):
        scale = weight.detach().abs().max().clamp_min(1e-8) / 127.0
        q = torch.round(weight.detach() / scale).clamp_(-128, 127).to(torch.int8)
        p = nn.Parameter(q, requires_grad=False)
        p.scale = scale
        return p

    def quantize_module(layer):
        for m in layer.modules():
            if isinstance(m, nn.Linear):
                m.weight = quantize_weight(m.weight)
        return layer

    for name, child in list(model.named_children()):
        if isinstance(child, orig_layer_impl):
            setattr(model, name, quantize_module(child))
        else:
            quantize_transformer_layer(orig_layer_impl, child, megatron, preln)
    return model


def quantize_module(layer, bits=8, groups=1):
    """MoQ-style in-place int8 quantisation of one transformer layer's four GEMM weights (qkv, attention output, the two
    MLP matrices): symmetric, ``groups`` groups per matrix (the MLP matrices use ``2*groups`` as in the reference);
    the weights stay in their float dtype holding the integer codes, scales are returned per matrix."""
    import torch
    scales = {}

    def q(t, g):
        flat = t.data.float().reshape(g, -1)
        s = (2**(bits - 1) - 1) / flat.abs().amax(dim=1, keepdim=True).clamp_min(1e-8)
        t.data = (flat * s).round().clamp(-2**(bits - 1), 2**(bits - 1) - 1).reshape(t.shape).to(t.dtype)
        return (1.0 / s).reshape(-1)

    for name, g in (("attn_qkvw", groups), ("attn_ow", groups), ("inter_w", 2 * groups), ("output_w", 2 * groups)):
        w = getattr(layer, name, None)
        if torch.is_tensor(w):
            scales[name] = q(w, g if w.numel() % g == 0 else 1)
    return scales
