"""Per-layer inference config (reference ``ops/transformer/inference/config.py``)."""
import json

import torch


class TransformerConfig:

    def __init__(self, hidden_size, intermediate_size, heads, num_hidden_layers):
        self.layer_id = -1
        self.hidden_size = hidden_size
        self.intermediate_size = intermediate_size
        self.heads = heads
        self.num_hidden_layers = num_hidden_layers


class DeepSpeedInferenceConfig(TransformerConfig):

    def __init__(self, hidden_size=-1, intermediate_size=-1, heads=-1, num_hidden_layers=-1, layer_norm_eps=1e-12,
                 local_rank=-1, mp_size=1, dtype=torch.float16, pre_layer_norm=True, norm_type="layer",
                 stochastic_mode=False, scale_attention=True, triangular_masking=True, local_attention=False, window_size=256,
                 rotary_dim=-1, rotate_half=False, rotate_every_two=True, return_tuple=True, mlp_after_attn=True,
                 mlp_act_func_type="gelu", training_mp_size=1, bigscience_bloom=False, max_out_tokens=1024,
                 min_out_tokens=1, enable_qkv_quantization=False, use_mup=False, scale_attn_by_inverse_layer_idx=False,
                 return_single_tuple=False, set_empty_params=False, transposed_mode=False, use_triton=False,
                 triton_autotune=False, num_kv=-1, rope_theta=10000, invert_mask=True):
        super().__init__(hidden_size, intermediate_size if intermediate_size > 0 else 4 * hidden_size, heads,
                         num_hidden_layers)
        self.dtype = dtype
        self.pre_layer_norm = pre_layer_norm
        self.norm_type = norm_type
        self.local_rank = local_rank
        self.stochastic_mode = stochastic_mode
        self.epsilon = layer_norm_eps
        self.mp_size = mp_size
        self.scale_attention = scale_attention
        self.triangular_masking = triangular_masking
        self.local_attention = local_attention
        self.window_size = window_size
        self.rotary_dim = rotary_dim
        self.rotate_half = rotate_half
        self.rotate_every_two = rotate_every_two
        self.return_tuple = return_tuple
        self.mlp_after_attn = mlp_after_attn
        self.mlp_act_func_type = mlp_act_func_type
        self.specialized_mode = False
        self.training_mp_size = training_mp_size
        self.bigscience_bloom = bigscience_bloom
        self.max_out_tokens = max_out_tokens
        self.min_out_tokens = min_out_tokens
        self.scale_attn_by_inverse_layer_idx = scale_attn_by_inverse_layer_idx
        self.enable_qkv_quantization = enable_qkv_quantization
        self.use_mup = use_mup
        self.return_single_tuple = return_single_tuple
        self.set_empty_params = set_empty_params
        self.transposed_mode = transposed_mode
        self.use_triton = use_triton
        self.triton_autotune = triton_autotune
        self.num_kv = num_kv
        self.rope_theta = rope_theta
        self.invert_mask = invert_mask

    @classmethod
    def from_dict(cls, json_object):
        config = DeepSpeedInferenceConfig()
        for key, value in json_object.items():
            config.__dict__[key] = value
        return config

    @classmethod
    def from_json_file(cls, json_file):
        with open(json_file, "r", encoding="utf-8") as reader:
            return cls.from_dict(json.loads(reader.read()))
