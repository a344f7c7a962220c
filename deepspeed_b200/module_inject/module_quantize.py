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
