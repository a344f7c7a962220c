"""AutoTP: discover the linears of a transformer and shard them column-/row-parallel.

Parity target: reference ``module_inject/auto_tp.py:192 AutoTP`` (+ ``fusedqkv_utils.py``, ``tp_shard.py``,
``runtime/tensor_parallel/tp_manager.py``).  Policy, like the reference: inside every repeated block, the
linears feeding the residual stream (names matching ``o_proj|out_proj|down_proj|dense_4h_to_h|c_proj|fc2|wo|w2``
or the last linear of the attention / MLP sub-module) become row-parallel ``LinearAllreduce``; all other
linears become column-parallel ``LinearLayer``; fused ``qkv`` / ``gate_up`` weights are split per part; the LM
head becomes ``LmHeadLinearAllreduce``; embeddings and norms stay replicated.  Attention modules get their
``num_heads`` style attributes divided by the TP degree.
"""
import re
from typing import Optional

import torch
from torch import nn

from deepspeed_b200 import comm as dist
from deepspeed_b200.utils import groups
from .layers import LinearAllreduce, LinearLayer, LmHeadLinearAllreduce

ROW_PATTERNS = re.compile(r"(o_proj|out_proj|down_proj|dense_4h_to_h|c_proj2?$|c_proj|fc2|wo$|w2$|attention\.dense|"
                          r"self_attention\.dense|out_lin|proj_out)")
FUSED_QKV = re.compile(r"(qkv_proj|query_key_value|c_attn|Wqkv|qkv$)")
FUSED_GATE_UP = re.compile(r"(gate_up_proj|w13)")
HEAD_ATTRS = ("num_heads", "num_attention_heads", "n_head", "n_heads", "num_key_value_heads", "num_kv_heads",
              "num_kv", "all_head_size", "embed_dim", "hidden_size", "split_size")


class ReplaceWithTensorSlicing:
    """Copy helpers that slice full checkpoints into this rank's TP shard (reference ``auto_tp.py:30``)."""

    def __init__(self, mp_group=None, mp_size=1, out_dim=1, in_dim=0):
        self.gpu_index = dist.get_rank(group=mp_group) if mp_group is not None else 0
        self.out_dim, self.in_dim, self.mp_size = out_dim, in_dim, mp_size

    def copy(self, dst, src, int8=False, allocate_tensor=False):
        if src is None:
            return src
        if dst.shape == src.shape:
            dst.data.copy_(src)
            return dst
        for dim in range(src.dim()):
            if dst.shape[dim] != src.shape[dim]:
                from .layers import shard_bounds
                s, e = shard_bounds(src.shape[dim], self.mp_size, self.gpu_index)
                dst.data.copy_(src.narrow(dim, s, e - s))
                return dst
        return dst

    def strided_copy(self, dst, src, num_splits, int8=False, allocate_tensor=False):
        from .layers import _split_rows
        dst.data.copy_(_split_rows(src, self.mp_size, self.gpu_index, fused_parts=num_splits))
        return dst


class AutoTP:

    def __init__(self, module, all_reduce_linears=None, prefix="", state_dict=None, linear_layer_setting=None,
                 orig_layer_impl=None, keep_module_on_host=False, mp_group=None, mp_size=None):
        self.module = module
        self.mp_group = mp_group
        self.mp_size = mp_size or (dist.get_world_size(mp_group) if mp_group is not None else 1)
        self.all_reduce_linears = set(all_reduce_linears or [])
        self.linear_policies = None

    @staticmethod
    def supported(model):
        return any(isinstance(m, nn.ModuleList) and len(m) > 1 for m in model.modules())

    @staticmethod
    def tp_parser(model):
        """List ``(block_class, [row-parallel linear names])`` like the reference's policy discovery."""
        out = []
        for m in model.modules():
            if isinstance(m, nn.ModuleList) and len(m) > 1:
                blk = m[0]
                rows = [n for n, l in blk.named_modules() if _is_linear(l) and ROW_PATTERNS.search(n)]
                out.append((type(blk), tuple(rows)))
        return out

    def _replace(self):
        n_col = n_row = 0
        for name, mod in list(self.module.named_modules()):
            for cname, child in list(mod.named_children()):
                full = f"{name}.{cname}" if name else cname
                if not _is_linear(child) or isinstance(child, (LinearLayer, LinearAllreduce)):
                    continue
                if not _inside_block(self.module, full) and not _is_lm_head(full):
                    continue
                if _is_lm_head(full):
                    if getattr(child, "weight", None) is not None and _tied_to_embedding(self.module, child):
                        continue  # tied heads stay replicated
                    setattr(mod, cname, LmHeadLinearAllreduce(_as_linear(child), self.mp_group, name=full))
                    n_row += 1
                elif ROW_PATTERNS.search(full) or cname in self.all_reduce_linears:
                    setattr(mod, cname, LinearAllreduce(_as_linear(child), self.mp_group, name=full))
                    n_row += 1
                else:
                    parts = 1
                    if FUSED_QKV.search(full):
                        parts = _qkv_parts(mod, child)
                    elif FUSED_GATE_UP.search(full):
                        parts = 2
                    setattr(mod, cname, LinearLayer(_as_linear(child), self.mp_group, fused_parts=parts, name=full))
                    n_col += 1
        self._fix_head_counts()
        return n_col, n_row

    def _fix_head_counts(self):
        for m in self.module.modules():
            has_tp_child = any(isinstance(c, (LinearLayer, LinearAllreduce)) for c in m.children())
            if not has_tp_child:
                continue
            for a in HEAD_ATTRS:
                v = getattr(m, a, None)
                if isinstance(v, int) and v >= self.mp_size and v % self.mp_size == 0 and a not in ("hidden_size", ):
                    setattr(m, a, v // self.mp_size)
            cfg = getattr(m, "cfg", None) or getattr(m, "config", None)
            if cfg is not None and not getattr(cfg, "_ds_tp_adjusted", False) and hasattr(cfg, "num_attention_heads"):
                import copy
                c2 = copy.copy(cfg)
                c2.num_attention_heads = cfg.num_attention_heads // self.mp_size
                if getattr(cfg, "num_key_value_heads", None):
                    c2.num_key_value_heads = max(1, cfg.num_key_value_heads // self.mp_size)
                c2._ds_tp_adjusted = True
                if hasattr(m, "cfg"):
                    m.cfg = c2
                else:
                    m.config = c2

    def replace(self):
        return self._replace()

    # ---- structure discovery helpers (reference ``auto_tp.py:215-282``) ------------------------------------------------
    @staticmethod
    def in_module_list(module, module_list):
        return any(type(item).__name__ == type(module).__name__ for item in module_list)

    @staticmethod
    def get_module_list(model):
        """One representative of every distinct block class found inside the model's ``ModuleList``s."""
        found = []
        for child in model.children():
            if isinstance(child, nn.ModuleList):
                for blk in child.children():
                    if not AutoTP.in_module_list(blk, found):
                        found.append(blk)
            else:
                found += [m for m in AutoTP.get_module_list(child) if not AutoTP.in_module_list(m, found)]
        return found

    @staticmethod
    def get_layers(parent, module):
        """Execution-ordered ``"parent.linear"`` names of a block, with ``"ln"`` markers where a layer norm sits."""
        out = []
        for key, sub in module._modules.items():
            if _is_linear(sub):
                out.append(f"{parent}.{key}")
            elif isinstance(sub, nn.LayerNorm) or key in ("LayerNorm", "layer_norm") or "RMSNorm" in type(sub).__name__:
                out.append("ln")
            elif sub is not None:
                out += AutoTP.get_layers(key, sub)
        return out

    @staticmethod
    def update_policy_list(policy_list, new_module, new_gems):
        """Merge ``(type(new_module), gems)`` into the policy list (gems of an already-listed class are unioned)."""
        for i, (cls, gems) in enumerate(policy_list):
            if cls is type(new_module):
                policy_list[i] = (cls, set(new_gems) | set(gems))
                return policy_list
        policy_list.append((type(new_module), new_gems))
        return policy_list

    @staticmethod
    def kernel_supported(module_list):
        """Does a kernel-injection policy exist for any of these block instances?"""
        from .replace_policy import replace_policies
        classes = set()
        for plcy in replace_policies:
            orig = getattr(plcy, "_orig_layer_class", None)
            for c in (orig if isinstance(orig, (list, tuple)) else [orig]):
                if c is not None:
                    classes.add(c)
        return any(type(m) in classes for m in module_list)

    def set_tensor_parallel_config(self, mp_size, mp_group):
        from .layers import is_autotp_training_mode
        if is_autotp_training_mode():
            from deepspeed_b200.utils import groups
            self.mp_group = groups.get_tensor_model_parallel_group()
            self.mp_size = groups.get_tensor_model_parallel_world_size()
            return
        self.mp_size, self.mp_group = mp_size, mp_group

    def update_mp_params(self, child):
        """Divide the head / width attributes of one attention module by the TP degree (once)."""
        if getattr(child, "replaced", False):
            return
        from .tp_shard import get_shard_size
        for a in HEAD_ATTRS:
            v = getattr(child, a, None)
            if isinstance(v, int):
                setattr(child, a, get_shard_size(v, self.mp_size))
        child.replaced = True

    def update_linear_policies(self):
        """Which leaf classes are sharded: ``nn.Linear`` (and HF ``Conv1D`` for GPT-2 style blocks)."""
        self.conv_linear_layer = False
        classes = [nn.Linear]
        try:
            from transformers.pytorch_utils import Conv1D
            classes.append(Conv1D)
        except ImportError:
            pass
        self.linear_policies = {c: self._replace for c in classes}
        return self.linear_policies

    @staticmethod
    def get_model_num_kv_heads(config):
        for name in ("multi_query_group_num", "num_kv_heads", "num_key_value_heads", "num_attention_heads", "n_heads",
                     "attention_heads"):
            v = getattr(config, name, None)
            if v is not None:
                return v
        return None


def _is_linear(m):
    return isinstance(m, nn.Linear) or (hasattr(m, "weight") and getattr(m, "in_features", None) is not None and
                                        isinstance(getattr(m, "weight", None), nn.Parameter) and m.weight.dim() == 2
                                        and type(m).__name__ in ("FlatLinear", "LMHead", "Conv1D"))


def _as_linear(m):
    if type(m).__name__ == "Conv1D":  # HF GPT-2: weight is [in, out]
        lin = nn.Linear(m.weight.shape[0], m.weight.shape[1], bias=m.bias is not None, device=m.weight.device,
                        dtype=m.weight.dtype)
        lin.weight.data = m.weight.data.t().contiguous()
        if m.bias is not None:
            lin.bias.data = m.bias.data
        return lin
    return m


def _inside_block(root, full_name):
    parts = full_name.split(".")
    mod = root
    for p in parts[:-1]:
        parent = mod
        mod = getattr(mod, p) if not p.isdigit() else mod[int(p)]
        if isinstance(parent, nn.ModuleList) and len(parent) > 1:
            return True
    return False


def _is_lm_head(name):
    return name.split(".")[-1] in ("lm_head", "embed_out", "output_layer")


def _tied_to_embedding(root, lin):
    return any(isinstance(m, nn.Embedding) and m.weight is lin.weight for m in root.modules())


def _qkv_parts(parent, lin):
    cfg = getattr(parent, "cfg", None) or getattr(parent, "config", None)
    if cfg is not None and getattr(cfg, "num_key_value_heads", None) and getattr(cfg, "num_attention_heads", None):
        hd = getattr(cfg, "head_dim", None) or cfg.hidden_size // cfg.num_attention_heads
        if cfg.num_key_value_heads != cfg.num_attention_heads:
            return [cfg.num_attention_heads * hd, cfg.num_key_value_heads * hd, cfg.num_key_value_heads * hd]
    return 3


def tp_model_init(model, tp_size, dtype, config=None, **kwargs):
    """Shard ``model`` in place for tensor-parallel training / inference (reference ``deepspeed.tp_model_init``)."""
    if not dist.is_initialized():
        dist.init_distributed()
    if tp_size <= 1:
        return model.to(dtype) if dtype is not None else model
    if groups.ranks_of("tp") is None:
        groups._init_tp_mesh_device(tensor_model_parallel_size=tp_size)
    tp_group = groups.get_tensor_model_parallel_group()
    if dtype is not None:
        model = model.to(dtype)
    AutoTP(model, mp_group=tp_group, mp_size=tp_size, **kwargs).replace()
    model.ds_autotp_parsed = True
    return model


def move(tensor, device, copy=True):
    from .layers import move as _move
    return _move(tensor, device, copy)


class Loading:
    """Checkpoint-to-module loading helpers for meta-initialised models (reference ``auto_tp.py:132``): modules that AutoTP
    does not shard (norms, embeddings, plain linears) still need their weights pulled from the state dict."""
    _LAYERS = (nn.Linear, nn.Embedding, nn.LayerNorm)
    _NAME_RX = ("RMSNorm", "LayerNorm", "RotaryEmbedding", "SharedEmbedding", "LearnedPositionalEmbedding", "FalconLinear",
                "MoEGate")

    @staticmethod
    def is_load_module(module):
        name = module._get_name()
        return isinstance(module, Loading._LAYERS) or any(k in name for k in Loading._NAME_RX)

    @staticmethod
    def _real(t):
        """A meta tensor cannot be copied into: swap it for host storage of the same shape."""
        if t.data.is_meta:
            return nn.Parameter(torch.empty_like(t.data, device="cpu"), requires_grad=t.requires_grad)
        return t

    @staticmethod
    def load_buffer(module, state_dict, prefix):
        for name, buf in list(module._buffers.items()):
            if buf is None:
                continue
            if buf.is_meta:
                module._buffers[name] = torch.empty_like(buf, device="cpu")
            if prefix + name in state_dict:
                module._buffers[name].data.copy_(state_dict[prefix + name])

    @staticmethod
    def load(module, state_dict, prefix, mp_group=None):
        slicer = ReplaceWithTensorSlicing(mp_group=mp_group)
        target = module if hasattr(module, "weight") else getattr(module, "norm", None)
        for attr in ("weight", "bias"):
            key = prefix + attr
            if key not in state_dict or target is None or getattr(target, attr, None) is None:
                continue
            cur = Loading._real(getattr(target, attr))
            if attr == "weight" and "query_key_value" in prefix and target is module:
                new = slicer.strided_copy(cur.data, state_dict[key], num_splits=3)
            else:
                new = slicer.copy(cur.data, state_dict[key])
            setattr(target, attr, new if isinstance(new, nn.Parameter) else nn.Parameter(new, requires_grad=cur.requires_grad))
