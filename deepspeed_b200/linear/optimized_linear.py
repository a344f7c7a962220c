"""``OptimizedLinear``: LoRA + base-weight sharding + quantised base weights behind one constructor
(reference ``linear/optimized_linear.py:18``)."""
import math

import torch
import torch.nn.functional as F
from torch import nn

from deepspeed_b200 import comm as dist
from .config import LoRAConfig, QuantizationConfig
from .quantization import QuantizedLinear, QuantizedParameter

_Linear = nn.Linear  # the stock class, immune to `linear.Init` patching nn.Linear


class OptimizedLinear(nn.Module):
    """Factory: plain ``nn.Linear`` (no configs), ``QuantizedLinear`` (quantisation only) or
    ``LoRAOptimizedLinear``."""

    def __new__(cls, input_dim: int, output_dim: int, lora_config: LoRAConfig = None,
                quantization_config: QuantizationConfig = None, dtype=torch.bfloat16, linear_cls=None, **kw):
        if lora_config is None and quantization_config is None:
            return _Linear(input_dim, output_dim, dtype=dtype, bias=kw.get("bias", False))
        if lora_config is not None:
            return LoRAOptimizedLinear(input_dim, output_dim, lora_config=lora_config,
                                       quantization_config=quantization_config, dtype=dtype)
        return QuantizedLinear(input_dim, output_dim, quantization_config=quantization_config, dtype=dtype)


class LoRAOptimizedLinear(nn.Module):

    def __init__(self, input_dim: int, output_dim: int, lora_config: LoRAConfig = None,
                 quantization_config: QuantizationConfig = None, device=None, dtype=torch.bfloat16):
        super().__init__()
        self.input_dim, self.output_dim = input_dim, output_dim
        self.lora_config, self.quantization_config = lora_config, quantization_config
        self.dtype = dtype
        self.sharding = max(1, lora_config.base_weight_sharding)
        self.zero_shards = self.sharding
        self.shard_rank = dist.get_rank() % self.sharding if (self.sharding > 1 and dist.is_initialized()) else 0
        assert input_dim % self.sharding == 0, "input_dim must be divisible by base_weight_sharding"
        self.sharded_weight_size = input_dim // self.sharding
        dev = device if device is not None else ("cuda" if torch.cuda.is_available() else "cpu")
        w = torch.empty(output_dim, self.sharded_weight_size, dtype=dtype, device=dev)
        nn.init.xavier_uniform_(w)
        if quantization_config is not None:
            self.weight = QuantizedParameter(w, quantization_config=quantization_config)
        else:
            self.weight = nn.Parameter(w, requires_grad=False)
        self.disabled = False
        self._shard_group = None
        if not lora_config.delay_lora_init:
            self.init_lora()

    def disable(self):
        self.disabled = True
        self.weight = nn.Parameter(self.weight.data, requires_grad=False)

    def init_lora(self):
        if self.disabled:
            return
        c = self.lora_config
        dev = self.weight.device
        self.lora_scaling_factor = c.lora_alpha / c.lora_r
        # lora_weight_1 is "A" (down), lora_weight_2 is "B" (up, zero-initialised -> adapter starts as identity)
        self.lora_weight_1 = _Linear(self.input_dim, c.lora_r, bias=False, dtype=self.dtype, device=dev)
        self.lora_weight_2 = _Linear(c.lora_r, self.output_dim, bias=False, dtype=self.dtype, device=dev)
        nn.init.kaiming_uniform_(self.lora_weight_1.weight, a=math.sqrt(5))
        nn.init.zeros_(self.lora_weight_2.weight)
        self.lora_weight_1.weight.requires_grad = True
        self.lora_weight_2.weight.requires_grad = True

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
        key = prefix + "weight"
        if key in state_dict:
            w = state_dict[key]
            if w.dim() == 2 and w.shape[1] == self.input_dim and self.sharding > 1:  # full weight: keep our columns
                w = w[:, self.shard_rank * self.sharded_weight_size:(self.shard_rank + 1) * self.sharded_weight_size]
            w = w.to(dtype=self.dtype, device=self.weight.device)
            if self.quantization_config is not None:
                self.weight = QuantizedParameter(w.contiguous(), quantization_config=self.quantization_config)
            else:
                self.weight = nn.Parameter(w.contiguous(), requires_grad=False)
            state_dict = {k: v for k, v in state_dict.items() if k != key}
        super()._load_from_state_dict(state_dict, prefix, local_metadata, False, missing_keys, unexpected_keys, error_msgs)
        if key in missing_keys:
            missing_keys.remove(key)

    def full_weight(self):
        base = self.weight.dequantized() if isinstance(self.weight, QuantizedParameter) else self.weight
        if self.sharding == 1 or not dist.is_initialized():
            return base
        if self._shard_group is None:
            world = dist.get_world_size()
            me = dist.get_rank()
            for s in range(0, world, self.sharding):
                ranks = list(range(s, min(s + self.sharding, world)))
                g = dist.new_group(ranks)
                if me in ranks:
                    self._shard_group = g
        parts = [torch.empty_like(base) for _ in range(self.sharding)]
        dist.all_gather(parts, base.contiguous(), group=self._shard_group)
        return torch.cat(parts, dim=1)

    def linear_without_F_linear(self, input, weight):
        return torch.matmul(input.reshape(-1, input.shape[-1]), weight.t()).view(*input.shape[:-1], weight.shape[0])

    def forward(self, input_tensor):
        base = self.full_weight()
        if self.disabled:
            return F.linear(input_tensor, base)
        with torch.no_grad():
            out = F.linear(input_tensor, base)
        lora = self.lora_weight_2(self.lora_weight_1(input_tensor))
        return out + self.lora_scaling_factor * lora
