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


# ---- SNIP-momentum block pruning (reference ``helper.py:257-322`` delegates to neural_compressor; native here) ----------
class SnipMomentumPruner:
    """Block-sparse magnitude×gradient pruning with momentum.

    Every optimizer step the criterion ``score ← β·score + |w ⊙ ∂L/∂w|`` (summed over ``N×M`` blocks) is updated; every
    ``pruning_frequency`` steps between ``start_step`` and ``end_step`` the mask is re-drawn so that the globally
    lowest-scoring blocks are removed, the sparsity following the cubic ramp ``s_t = s·(1-(1-p)^3)``.  After every
    optimizer step the mask is re-applied, so pruned blocks stay zero."""

    def __init__(self, modules, target_sparsity, pattern="4x1", pruning_frequency=1, start_step=0, end_step=0, beta=0.9):
        import torch
        self.modules = dict(modules)
        self.target_sparsity = float(target_sparsity)
        n, m = (int(v) for v in str(pattern).lower().split("x"))
        self.block = (n, m)
        self.pruning_frequency = max(1, int(pruning_frequency))
        self.start_step, self.end_step, self.beta = int(start_step), max(int(end_step), int(start_step)), beta
        self.global_step = 0
        self.scores = {k: torch.zeros(self._blocks(mod.weight), device=mod.weight.device) for k, mod in self.modules.items()}
        self.masks = {k: torch.ones_like(mod.weight, dtype=torch.bool) for k, mod in self.modules.items()}
        self.current_sparsity = 0.0

    def _blocks(self, w):
        n, m = self.block
        assert w.dim() == 2 and w.shape[0] % n == 0 and w.shape[1] % m == 0, f"weight {tuple(w.shape)} not tileable by {n}x{m}"
        return w.shape[0] // n, w.shape[1] // m

    def _reduce(self, t):
        n, m = self.block
        r, c = t.shape[0] // n, t.shape[1] // m
        return t.reshape(r, n, c, m).sum(dim=(1, 3))

    def _expand(self, blk):
        n, m = self.block
        return blk.repeat_interleave(n, 0).repeat_interleave(m, 1)

    def _sparsity_at(self, step):
        if step >= self.end_step:
            return self.target_sparsity
        span = max(1, self.end_step - self.start_step)
        p = min(1.0, max(0.0, (step - self.start_step) / span))
        return self.target_sparsity * (1.0 - (1.0 - p)**3)

    # ---- hooks
    def on_step_begin(self, local_step=0):
        import torch
        step = self.global_step
        if step < self.start_step or step > self.end_step or (step - self.start_step) % self.pruning_frequency:
            return
        want = self._sparsity_at(step)
        allv = torch.cat([s.reshape(-1) for s in self.scores.values()])
        k = int(want * allv.numel())
        if k <= 0 or not bool(allv.any()):
            return
        thr = torch.kthvalue(allv, k).values
        for name, s in self.scores.items():
            self.masks[name] = self._expand(s > thr)
        self.current_sparsity = want
        self._apply()

    def on_before_optimizer_step(self):
        import torch
        with torch.no_grad():
            for name, mod in self.modules.items():
                if mod.weight.grad is not None:
                    self.scores[name].mul_(self.beta).add_(self._reduce((mod.weight * mod.weight.grad).abs().float()))

    def on_after_optimizer_step(self):
        self._apply()
        self.global_step += 1

    def _apply(self):
        import torch
        with torch.no_grad():
            for name, mod in self.modules.items():
                mod.weight.mul_(self.masks[name].to(mod.weight.dtype))

    def sparsity(self):
        tot = sum(m.numel() for m in self.masks.values())
        return 1.0 - sum(int(m.sum()) for m in self.masks.values()) / max(1, tot)


def generate_pruners(config, model):
    """``config``: dict (or object) with ``target_sparsity, pattern, pruning_frequency, start_step, end_step,
    excluded_op_names``.  One pruner over every 2-D ``nn.Linear`` weight not excluded and tileable by the block pattern."""
    import re
    import torch
    get = (lambda k, d=None: config.get(k, d)) if isinstance(config, dict) else (lambda k, d=None: getattr(config, k, d))
    excluded = list(get("excluded_op_names", []) or [])
    n, m = (int(v) for v in str(get("pattern", "4x1")).lower().split("x"))
    mods = {}
    for name, mod in model.named_modules():
        if isinstance(mod, torch.nn.Linear) and not any(re.search(p, name) for p in excluded):
            if mod.weight.shape[0] % n == 0 and mod.weight.shape[1] % m == 0:
                mods[name] = mod
    if not mods:
        from deepspeed_b200.utils.logging import logger
        logger.warning("one pruner hooks no layers, please have a check")
    return [SnipMomentumPruner(mods, get("target_sparsity", 0.9), get("pattern", "4x1"), get("pruning_frequency", 1),
                               get("start_step", 0), get("end_step", 0))]


def register_on_step_begin(model):
    """Forward pre-hook that lets every pruner of ``model.pruners`` refresh its mask at the start of a step."""

    def hook(module, inputs):
        if module.training:
            for pruner in module.pruners:
                pruner.on_step_begin(0)

    return model.register_forward_pre_hook(hook)


def rewrite_optimizer_step(opt):
    """Wrap ``opt.step`` with the pruners' before / after callbacks (``opt.pruners`` is read at call time)."""
    import types

    def new_step(self, closure=None):
        for pruner in getattr(self, "pruners", ()):
            pruner.on_before_optimizer_step()
        res = self.orig_step(closure) if closure is not None else self.orig_step()
        for pruner in getattr(self, "pruners", ()):
            pruner.on_after_optimizer_step()
        return res

    opt.orig_step = opt.step
    opt.step = types.MethodType(new_step, opt)
    return opt
