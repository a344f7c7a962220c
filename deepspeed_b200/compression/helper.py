"""Module replacement helpers (reference ``compression/helper.py``)."""
import torch
from torch import nn

from . import constants as C
from .basic_layer import (BNLayer_Compress, ColumnParallelLinear_Compress, Conv2dLayer_Compress, Embedding_Compress,
                          LinearLayer_Compress, RowParallelLinear_Compress)


def recursive_getattr(model, module_name):
    out = model
    for name in module_name.split("."):
        out = getattr(out, name)
    return out


def recursive_setattr(model, module_name, module):
    parts = module_name.split(".")
    parent = recursive_getattr(model, ".".join(parts[:-1])) if len(parts) > 1 else model
    setattr(parent, parts[-1], module)


def is_module_compressible(module, mpu=None):
    ok = isinstance(module, (nn.Linear, nn.Conv2d, nn.Embedding, nn.BatchNorm2d))
    if mpu is not None and not ok:
        ok = isinstance(module, (getattr(mpu, "RowParallelLinear", ()), getattr(mpu, "ColumnParallelLinear", ())))
    return ok


def _to_compress_layer(old, mpu=None):
    if isinstance(old, (LinearLayer_Compress, Conv2dLayer_Compress, Embedding_Compress, BNLayer_Compress)):
        return old
    dev, dt = old.weight.device, old.weight.dtype
    if isinstance(old, nn.Linear):
        new = LinearLayer_Compress(old.in_features, old.out_features, bias=old.bias is not None)
    elif isinstance(old, nn.Conv2d):
        new = Conv2dLayer_Compress(old.in_channels, old.out_channels, old.kernel_size, old.stride, old.padding, old.dilation,
                                   old.groups, old.bias is not None, old.padding_mode)
    elif isinstance(old, nn.BatchNorm2d):
        new = BNLayer_Compress(old.num_features, old.eps, old.momentum, old.affine, old.track_running_stats)
        new.load_state_dict(old.state_dict())
        return new.to(device=dev, dtype=dt)
    elif isinstance(old, nn.Embedding):
        new = Embedding_Compress(old.num_embeddings, old.embedding_dim, old.padding_idx, old.max_norm, old.norm_type,
                                 old.scale_grad_by_freq, old.sparse)
    elif mpu is not None and isinstance(old, getattr(mpu, "ColumnParallelLinear", ())):
        new = ColumnParallelLinear_Compress(mpu, old.input_size, old.output_size, bias=old.bias is not None,
                                            gather_output=old.gather_output, skip_bias_add=old.skip_bias_add)
    elif mpu is not None and isinstance(old, getattr(mpu, "RowParallelLinear", ())):
        new = RowParallelLinear_Compress(mpu, old.input_size, old.output_size, bias=old.bias is not None,
                                         input_is_parallel=old.input_is_parallel, skip_bias_add=old.skip_bias_add)
    else:
        return None
    new = new.to(device=dev, dtype=dt)
    new.weight.data = old.weight.data
    if getattr(old, "bias", None) is not None:
        new.bias.data = old.bias.data
    return new


def module_replacement(model, module_name, compression_technique=None, mpu=None):
    """Swap ``module_name`` for its compressible counterpart and switch on the given techniques
    (``{technique: {shared..., group params...}}``)."""
    old = recursive_getattr(model, module_name)
    new = _to_compress_layer(old, mpu)
    if new is None:
        return old
    for k, v in (compression_technique or {}).items():
        if k == C.SPARSE_PRUNING:
            if v[C.TECHNIQUE_ENABLED]:
                new.enable_sparse_pruning(v[C.SPARSE_PRUNING_DENSE_RATIO], v[C.SPARSE_PRUNING_METHOD])
        elif k == C.ROW_PRUNING:
            if v[C.TECHNIQUE_ENABLED]:
                new.enable_row_pruning(v[C.ROW_PRUNING_DENSE_RATIO], v[C.ROW_PRUNING_METHOD])
        elif k == C.HEAD_PRUNING:
            if v[C.TECHNIQUE_ENABLED]:
                new.enable_head_pruning(v[C.HEAD_PRUNING_DENSE_RATIO], v[C.HEAD_PRUNING_METHOD], v[C.HEAD_PRUNING_NUM_HEADS])
        elif k == C.CHANNEL_PRUNING:
            if v[C.TECHNIQUE_ENABLED]:
                new.enable_channel_pruning(v[C.CHANNEL_PRUNING_DENSE_RATIO], v[C.CHANNEL_PRUNING_METHOD])
        elif k == C.ACTIVATION_QUANTIZATION:
            if v[C.TECHNIQUE_ENABLED]:
                new.enable_activation_quantization(v[C.ACTIVATION_QUANTIZE_BITS], v[C.ACTIVATION_QUANTIZE_TYPE],
                                                   v[C.ACTIVATION_QUANTIZE_RANGE])
        elif k == C.WEIGHT_QUANTIZATION:
            if v[C.TECHNIQUE_ENABLED]:
                new.enable_weight_quantization(v[C.WEIGHT_QUANTIZE_START_BITS], v[C.WEIGHT_QUANTIZE_TARGET_BITS],
                                               v[C.WEIGHT_QUANTIZATION_PERIOD], v[C.WEIGHT_QUANTIZE_IN_FORWARD_ENABLED],
                                               v[C.WEIGHT_QUANTIZE_TYPE], v[C.WEIGHT_QUANTIZE_GROUPS])
        else:
            raise NotImplementedError(f"Compression technique {k} is not implemented")
    recursive_setattr(model, module_name, new)
    return new


def compression_preparation(model, compression_technique_list, mpu):
    """Phase 1: make every compressible module a ``*_Compress`` layer; phase 2: enable techniques per group."""
    for name, module in list(model.named_modules()):
        if name and is_module_compressible(module, mpu):
            module_replacement(model, name, mpu=mpu)
    for module_name_lists, _, technique in compression_technique_list:
        for names in module_name_lists:
            for name in names:
                module_replacement(model, name, technique)
    return model


def fix_compression(model, module_name, compression_technique, mask=None, dim_reduction=False):
    """Bake the (scheduled) techniques of ``module_name`` into its weights; returns the structural mask so the
    related (next / producing) modules can be resized."""
    module = recursive_getattr(model, module_name)
    for k, v in compression_technique.items():
        if k == C.WEIGHT_QUANTIZATION and v[C.WEIGHT_QUANTIZE_IN_FORWARD_ENABLED] and v[C.TECHNIQUE_ENABLED]:
            return module.fix_weight_quantization()
        if k == C.SPARSE_PRUNING and v[C.TECHNIQUE_ENABLED]:
            return module.fix_sparse_pruning_helper()
        if k == C.ROW_PRUNING and (v[C.TECHNIQUE_ENABLED] or mask is not None):
            return module.fix_row_col_pruning_helper(mask, dim_reduction=dim_reduction)
        if k == C.HEAD_PRUNING and (v[C.TECHNIQUE_ENABLED] or mask is not None):
            return module.fix_head_pruning_helper(mask, v[C.HEAD_PRUNING_NUM_HEADS], dim_reduction=dim_reduction)
        if k == C.CHANNEL_PRUNING and (v[C.TECHNIQUE_ENABLED] or mask is not None):
            return module.fix_channel_pruning_helper(mask, dim_reduction=dim_reduction)
    return None


def convert_conv1d_to_linear(model, convert_type):
    """HF GPT-2 ``Conv1D`` -> ``nn.Linear`` (so compression applies)."""
    if hasattr(model, "module"):
        model = model.module
    for name, module in list(model.named_modules()):
        if name and isinstance(module, convert_type):
            lin = nn.Linear(module.weight.shape[0], module.weight.shape[1], bias=module.bias is not None)
            lin.weight.data = module.weight.data.t().contiguous()
            if module.bias is not None:
                lin.bias.data = module.bias.data
            recursive_setattr(model, name, lin.to(module.weight.device))
    return model
