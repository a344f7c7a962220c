"""Per-layer inference config (reference ``ops/transformer/inference/config.py``)."""
import json

import torch

# field -> default; ``DeepSpeedInferenceConfig(**overrides)`` accepts exactly these names (plus the four positional
# architecture sizes).  ``layer_norm_eps`` is stored as ``epsilon``.
_DEFAULTS = dict(
    layer_norm_eps=1e-12, local_rank=-1, mp_size=1, dtype=torch.float16, pre_layer_norm=True, norm_type="layer",
    stochastic_mode=False, scale_attention=True, triangular_masking=True, local_attention=False, window_size=256, rotary_dim=-1,
    rotate_half=False, rotate_every_two=True, return_tuple=True, mlp_after_attn=True, mlp_act_func_type="gelu", training_mp_size=1,
    bigscience_bloom=False, max_out_tokens=1024, min_out_tokens=1, enable_qkv_quantization=False, use_mup=False,
    scale_attn_by_inverse_layer_idx=False, return_single_tuple=False, set_empty_params=False, transposed_mode=False,
    use_triton=False, triton_autotune=False, num_kv=-1, rope_theta=10000, invert_mask=True)
_RENAMED = {"layer_norm_eps": "epsilon"}


class TransformerConfig:

    def __init__(self, hidden_size, intermediate_size, heads, num_hidden_layers):
        self.layer_id = -1
        self.hidden_size, self.intermediate_size = hidden_size, intermediate_size
        self.heads, self.num_hidden_layers = heads, num_hidden_layers


class DeepSpeedInferenceConfig(TransformerConfig):

    def __init__(self, hidden_size=-1, intermediate_size=-1, heads=-1, num_hidden_layers=-1, *args, **overrides):
        super().__init__(hidden_size, intermediate_size if intermediate_size > 0 else 4 * hidden_size, heads, num_hidden_layers)
        names = list(_DEFAULTS)
        if len(args) > len(names):
            raise TypeError(f"too many positional arguments ({len(args) + 4})")
        overrides = {**dict(zip(names, args)), **overrides}  # the reference signature is positional-friendly
        unknown = set(overrides) - set(names)
        if unknown:
            raise TypeError(f"unknown DeepSpeedInferenceConfig field(s): {sorted(unknown)}")
        for name, default in _DEFAULTS.items():
            setattr(self, _RENAMED.get(name, name), overrides.get(name, default))
        self.specialized_mode = False

    @classmethod
    def from_dict(cls, json_object):
        config = cls()
        config.__dict__.update(json_object)
        return config

    @classmethod
    def from_json_file(cls, json_file):
        with open(json_file, "r", encoding="utf-8") as reader:
            return cls.from_dict(json.loads(reader.read()))
