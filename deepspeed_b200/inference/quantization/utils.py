"""Group-wise asymmetric weight quantizer / dequantizer objects for ZeroQuant-style inference
(reference ``inference/quantization/utils.py``: ``Quantizer``, ``DeQuantizer``, helpers).

Storage convention (kept from the reference so checkpoints interchange): unsigned codes ``round((x - min) * scale)`` with
``scale = (2^bits - 1) / (max - min)`` per group along ``group_dim``; 4-bit codes are packed two per byte along the LAST
dimension, even element in the high nibble.
"""
from typing import Dict, Tuple

import torch
from torch import Tensor, nn


def tensor_clamp(tensor: Tensor, lo, hi) -> Tensor:
    return tensor.clamp_(lo, hi)


def tensor_round(tensor: Tensor) -> Tensor:
    return tensor.round_()


def _check(config):
    assert config["num_bits"] in (4, 8), "Only INT4 and INT8 quantization is supported."
    assert config["symmetric"] is False, "Only asymmetric quantization is supported at this moment."


def _grouped(t, config):
    d, gs = config["group_dim"] % t.dim(), config["group_size"]
    assert t.shape[d] % gs == 0, f"Tensor shape: {tuple(t.shape)} quantization config {config}"
    return t.reshape(*t.shape[:d], t.shape[d] // gs, gs, *t.shape[d + 1:]), d + 1


class Quantizer:

    def __init__(self, config: Dict) -> None:
        _check(config)
        self.config = config

    def quantize(self, tensor: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
        """-> (codes uint8 [packed for 4 bit], scale, min) with scale/min shaped like the grouped tensor (group axis 1)."""
        shape = tensor.shape
        g, axis = _grouped(tensor.detach().float(), self.config)
        q_range = 2**self.config["num_bits"] - 1
        lo, hi = g.amin(dim=axis, keepdim=True), g.amax(dim=axis, keepdim=True)
        scale = q_range / (hi - lo).clamp_min(1e-10)
        codes = tensor_round(tensor_clamp((g - lo) * scale, 0, q_range)).to(torch.uint8).reshape(shape)
        if self.config["num_bits"] == 4:
            assert shape[-1] % 2 == 0
            codes = (codes[..., 0::2] << 4) | codes[..., 1::2]
        return codes, scale.to(tensor.dtype), lo.to(tensor.dtype)


class DeQuantizer:

    def __init__(self, config: Dict, dtype: torch.dtype) -> None:
        _check(config)
        self.config, self.dtype = config, dtype

    def dequantize(self, tensor: Tensor, quant_scale: Tensor, quant_min: Tensor) -> Tensor:
        assert tensor.dtype == torch.uint8
        if self.config["num_bits"] == 4:
            tensor = torch.stack((tensor >> 4, tensor & 0xF), dim=-1).reshape(*tensor.shape[:-1], tensor.shape[-1] * 2)
        shape = tensor.shape
        g, _ = _grouped(tensor, self.config)
        return (g.to(self.dtype) / quant_scale.to(self.dtype) + quant_min.to(self.dtype)).reshape(shape)


def recursive_setattr(model, module_name, module):
    """``setattr`` along a dotted path (``"a.b.3.c"``)."""
    head, _, tail = module_name.partition(".")
    if tail:
        recursive_setattr(getattr(model, head), tail, module)
    else:
        setattr(model, head, module)


def concat_to_compat_param(quantized_weight: Tensor, quant_scale: Tensor, quant_min: Tensor, return_param: bool = True):
    """Pack codes + per-group scale/min into ONE flat byte tensor (so ZeRO-3 can partition / gather it like any other
    parameter); the inverse is :func:`split_compat_param`."""
    dt = quant_scale.dtype
    blob = torch.cat([quantized_weight.reshape(-1).view(torch.uint8), quant_scale.contiguous().reshape(-1).view(torch.uint8),
                      quant_min.to(dt).contiguous().reshape(-1).view(torch.uint8)])
    return nn.Parameter(blob, requires_grad=False) if return_param else blob


def split_compat_param(blob: Tensor, weight_shape, n_groups, dtype):
    nq = 1
    for s in weight_shape:
        nq *= s
    es = torch.empty(0, dtype=dtype).element_size()
    codes = blob[:nq].reshape(weight_shape)
    scale = blob[nq:nq + n_groups * es].view(dtype)
    mn = blob[nq + n_groups * es:nq + 2 * n_groups * es].view(dtype)
    return codes, scale, mn


def _quantize_param(param: nn.Parameter, quant_config: Dict):
    """Replace ``param.data`` by its packed quantized form and remember how to undo it."""
    assert not getattr(param, "weight_quantized", False), "Parameter has already been quantized."
    codes, scale, mn = Quantizer(quant_config).quantize(param.data)
    param.quant_shape, param.quant_groups, param.quant_dtype = tuple(codes.shape), scale.numel(), param.dtype
    param.quant_scale_shape = tuple(scale.shape)
    param.quant_config = quant_config
    param.requires_grad = False
    param.data = concat_to_compat_param(codes, scale, mn, return_param=False)
    param.weight_quantized = True


def dequantize_param(param: nn.Parameter) -> Tensor:
    codes, scale, mn = split_compat_param(param.data, param.quant_shape, param.quant_groups, param.quant_dtype)
    return DeQuantizer(param.quant_config, param.quant_dtype).dequantize(codes, scale.view(param.quant_scale_shape),
                                                                          mn.view(param.quant_scale_shape))


def wrap_quantized_functional(f):
    """Decorator for ``F.linear`` / ``F.embedding``-like functions: a quantized weight argument is dequantized first."""
    import functools

    @functools.wraps(f)
    def wrapper(input, weight, *args, **kwargs):
        if getattr(weight, "weight_quantized", False):
            weight = dequantize_param(weight)
        return f(input, weight, *args, **kwargs)

    return wrapper


def wrap_load_from_state_dict(f):
    """Decorator for ``Module._load_from_state_dict``: after loading, (re)quantize parameters marked for quantization."""
    import functools

    @functools.wraps(f)
    def wrapper(module, state_dict, prefix, *args, **kwargs):
        marked = {n: p for n, p in module._parameters.items() if p is not None and getattr(p, "weight_quantized", False)}
        for n, p in marked.items():  # let the float tensor load into a float slot
            p.data = torch.empty(p.quant_full_shape if hasattr(p, "quant_full_shape") else (0, ), dtype=p.quant_dtype)
            p.weight_quantized = False
        f(module, state_dict, prefix, *args, **kwargs)
        for n, p in marked.items():
            if p.numel() > 0:
                _quantize_param(p, p.quant_config)

    return wrapper


def get_quantizer_module():
    """The native quantiser op (reference ``utils.py:get_quantizer_cuda_module``)."""
    from deepspeed_b200.ops.quantizer import quantizer
    return quantizer


get_quantizer_cuda_module = get_quantizer_module


def get_AsyncPartitionedParameterSwapper(model: nn.Module):
    """The NVMe parameter swapper attached to any ZeRO-3 parameter of ``model`` (``None`` when parameters stay resident)."""
    for p in model.parameters():
        sw = getattr(p, "nvme_swapper", None)
        if sw is not None:
            return sw
        zo = getattr(p, "_ds_zero", None)
        zo = zo() if callable(zo) else None
        if zo is not None and getattr(zo, "param_swapper", None) is not None:
            return zo.param_swapper
    return None
