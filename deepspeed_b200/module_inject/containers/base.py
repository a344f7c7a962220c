"""Container = policy + fused layer construction + tensor-parallel slicing (reference
``module_inject/containers/base.py:BaseTransformerContainer``).

``policy`` reads the original layer; the container builds a ``DeepSpeedInferenceConfig`` from the policy's facts, slices
the canonical tensors for this tensor-parallel rank, instantiates the family's fused layer, copies the weights in and
wraps it in an adapter that speaks the original layer's calling convention (argument names, tuple-or-tensor return,
HF ``Cache`` bookkeeping) so the surrounding Hugging Face model code keeps working unchanged.
"""
from abc import ABC
import inspect

import torch
from torch import nn

from deepspeed_b200.ops.transformer.inference.config import DeepSpeedInferenceConfig
from deepspeed_b200.utils.types import ActivationFuncType, NormType

_ACT_NAMES = {ActivationFuncType.GELU: "gelu", ActivationFuncType.ReLU: "relu", ActivationFuncType.GATED_GELU: "gated_gelu",
              ActivationFuncType.GATED_SILU: "gated_silu"}


def _returns_tuple(layer):
    """Does the original layer's forward end in ``return a, b`` / ``return outputs`` (tuple) or ``return hidden`` (tensor)?"""
    try:
        src = inspect.getsource(type(layer).forward)
    except (OSError, TypeError):
        return True
    rets = [ln.split("#")[0].strip() for ln in src.splitlines() if ln.strip().startswith("return")]
    if not rets:
        return True
    last = rets[-1][len("return"):].strip()
    return "," in last or last.startswith("(") or last == "outputs"


class InjectedLayer(nn.Module):
    """Adapter around a fused layer that mimics the replaced layer's signature."""

    def __init__(self, fused, orig_layer, layer_idx, causal):
        super().__init__()
        self.fused = fused
        self.layer_idx = layer_idx
        self.causal = causal
        self._tuple = _returns_tuple(orig_layer)
        self._params = [p for p in inspect.signature(type(orig_layer).forward).parameters if p != "self"]

    def forward(self, *args, **kwargs):
        bound = dict(zip(self._params, args))
        bound.update(kwargs)
        x = bound.get("hidden_states", args[0] if args else None)
        if isinstance(x, tuple):
            x = x[0]
        mask = bound.get("attention_mask")
        cache = bound.get("past_key_values", bound.get("layer_past", bound.get("past_key_value")))
        use_cache = bool(bound.get("use_cache")) or cache is not None
        if self.causal and cache is not None and hasattr(cache, "update") and hasattr(cache, "get_seq_length"):
            # the fused layer owns the real KV cache; advance the HF cache's length with a 1-element placeholder so the
            # surrounding generate() loop slices input_ids / builds positions correctly
            if cache.get_seq_length(self.layer_idx) == 0:
                self.fused.reset_cache()
            ph = x.new_zeros(x.shape[0], 1, x.shape[1], 1)
            cache.update(ph, ph, self.layer_idx)
        elif not use_cache:
            self.fused.reset_cache()
        out = self.fused(x, attention_mask=mask, use_cache=use_cache)
        out = out[0] if isinstance(out, tuple) else out
        return (out, None) if self._tuple else out


class BaseTransformerContainer:
    """Subclasses set ``layer_class`` (a ``DeepSpeed*Inference``); everything else derives from the policy."""
    layer_class = None

    def __init__(self, policy, config=None, model_config=None, layer_id=0, child=None):
        self.policy = policy
        self.config = config
        self.model_config = model_config
        self.layer_id = layer_id
        self.child = child if child is not None else getattr(policy, "client_module", None)
        self.mp_size = getattr(getattr(config, "tensor_parallel", None), "tp_size", 1) if config is not None else 1
        self.mp_group = None
        self.dtype = getattr(config, "dtype", None) or torch.float16
        if self.dtype == torch.int8:
            self.dtype = torch.float16
        self.max_out_tokens = getattr(config, "max_out_tokens", 1024) if config is not None else 1024
        self.hidden_size, self.num_attention_heads, self.layernorm_epsilon, self.intermediate_size = policy.get_hidden_heads()
        if not self.intermediate_size or self.intermediate_size < 0:
            self.intermediate_size = 4 * self.hidden_size
        self.module = None
        self.ds_model_config = None

    # ---- config ---------------------------------------------------------------------------------------------------
    def create_ds_model_config(self):
        p = self.policy
        rot_dim, rot_half, theta = p.rotary()
        window = p.local_window()
        self.ds_model_config = DeepSpeedInferenceConfig(
            hidden_size=self.hidden_size, intermediate_size=self.intermediate_size, heads=self.num_attention_heads,
            layer_norm_eps=self.layernorm_epsilon, dtype=self.dtype, pre_layer_norm=p.pre_attn_norm,
            norm_type="rms" if p.norm_type == NormType.RMSNorm else "layer", mp_size=self.mp_size,
            scale_attention=p.scale_attention, triangular_masking=p.causal(), local_attention=window > 0,
            window_size=window or 256, rotary_dim=rot_dim, rotate_half=rot_half, rotate_every_two=not rot_half,
            return_tuple=False, mlp_after_attn=p.mlp_after_attn(),
            mlp_act_func_type=getattr(p, "act_name", None) or _ACT_NAMES.get(p.mlp_act_func_type, "gelu"),
            bigscience_bloom=p.uses_alibi(), max_out_tokens=self.max_out_tokens, use_mup=p.use_mup, num_kv=p.num_kv_heads(),
            rope_theta=theta)
        self.ds_model_config.parallel_mlp_own_norm = p.parallel_mlp_own_norm()
        return self.ds_model_config

    def set_tensor_parallel_config(self, mp_size, mp_group):
        self.mp_size, self.mp_group = mp_size, mp_group

    # ---- tensors ----------------------------------------------------------------------------------------------------
    def initialize_tensors(self, enable_training=False):
        self.qkvw, self.qkvb, self.dense_w, self.dense_b = self.policy.attention()
        self._h4h_w, self._h4h_b, self._4hh_w, self._4hh_b = self.policy.mlp()
        self.attn_nw, self.attn_nb, self.input_nw, self.input_nb = self.policy.layernorm()

    def _rank(self):
        if self.mp_group is None or self.mp_size <= 1:
            return 0
        from deepspeed_b200 import comm as dist
        return dist.get_rank(self.mp_group)

    def apply_tensor_parallelism(self, mp_replace=None):
        """Column-split QKV (by head) and the first MLP GEMM, row-split the two output projections."""
        tp, r = self.mp_size, self._rank()
        if tp <= 1:
            return
        heads = self.num_attention_heads
        kv = self.policy.num_kv_heads()
        kv = kv if kv and kv > 0 else heads
        d = self.hidden_size // heads
        assert heads % tp == 0 and kv % tp == 0, f"heads ({heads}/{kv}) must divide over tp={tp}"

        def split_qkv(t):
            if t is None:
                return None
            q, k, v = t[:heads * d], t[heads * d:(heads + kv) * d], t[(heads + kv) * d:]
            take = lambda x, n: x[r * (n // tp) * d:(r + 1) * (n // tp) * d]
            return torch.cat([take(q, heads), take(k, kv), take(v, kv)], dim=0)

        def rows(t, parts=1):
            if t is None:
                return None
            chunks = t.chunk(parts, dim=0)  # gated MLP: gate and up are split independently
            return torch.cat([c.chunk(tp, dim=0)[r] for c in chunks], dim=0)

        gated = self._h4h_w.shape[0] == 2 * self.intermediate_size
        self.qkvw, self.qkvb = split_qkv(self.qkvw), split_qkv(self.qkvb)
        self.dense_w = self.dense_w.chunk(tp, dim=1)[r]
        self._h4h_w, self._h4h_b = rows(self._h4h_w, 2 if gated else 1), rows(self._h4h_b, 2 if gated else 1)
        self._4hh_w = self._4hh_w.chunk(tp, dim=1)[r]
        if r != 0:  # biases of row-parallel projections are added once (after the all-reduce) -> keep them on rank 0
            self.dense_b = None if self.dense_b is None else torch.zeros_like(self.dense_b)
            self._4hh_b = None if self._4hh_b is None else torch.zeros_like(self._4hh_b)

    # ---- module -------------------------------------------------------------------------------------------------------
    def create_module(self, config=None):
        cfg = config or self.ds_model_config or self.create_ds_model_config()
        self.module = self.layer_class(cfg, mp_group=self.mp_group)
        self.module.config.layer_id = self.layer_id
        return self.module

    def copy_data_to_new_module(self):
        m = self.module

        def put(dst, src):
            with torch.no_grad():
                if src is None:
                    dst.zero_()
                else:
                    dst.copy_(src.detach().to(dst.dtype).reshape(dst.shape))

        put(m.attn_qkvw, self.qkvw), put(m.attn_qkvb, self.qkvb)
        put(m.attn_ow, self.dense_w), put(m.attn_ob, self.dense_b)
        put(m.inter_w, self._h4h_w), put(m.inter_b, self._h4h_b)
        put(m.output_w, self._4hh_w), put(m.output_b, self._4hh_b)
        put(m.attn_nw, self.attn_nw), put(m.attn_nb, self.attn_nb)
        put(m.norm_w, self.input_nw), put(m.norm_b, self.input_nb)
        if self.policy.norm_type != NormType.RMSNorm:
            return
        # RMSNorm has no bias; the zeros written above are simply unused

    def build(self, device=None):
        """config -> tensors -> TP slicing -> fused layer with weights, wrapped for drop-in use."""
        self.create_ds_model_config()
        self.initialize_tensors()
        self.apply_tensor_parallelism()
        self.create_module()
        self.copy_data_to_new_module()
        if device is not None:
            self.module.to(device)
        return InjectedLayer(self.module, self.child, self.layer_id, self.policy.causal())


class BaseConvolutionContainer(ABC):
    """Placeholder base for convolutional (diffusers) containers: they need no shared state (reference ``base.py:20``)."""

    def __init__(self):
        pass
