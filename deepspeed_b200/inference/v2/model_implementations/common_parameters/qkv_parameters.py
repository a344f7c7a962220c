"""Reference ``model_implementations/common_parameters/qkv_parameters.py``."""
import torch

from ..parameter_base import ParameterBase, ParamList  # noqa: F401


class FusedQKVParameter(ParameterBase):
    """Checkpoint already stores [q | k | v]."""
    params: torch.Tensor

    def finalize(self) -> torch.Tensor:
        return self.inference_model.transform_qkv_param(self.params)


class UnfusedQKVParameter(ParameterBase):
    """Separate q / k / v projections (MHA), fused by concatenation."""
    q_params: torch.Tensor
    k_params: torch.Tensor
    v_params: torch.Tensor

    def finalize(self) -> torch.Tensor:
        return self.inference_model.transform_qkv_param(torch.cat([self.q_params, self.k_params, self.v_params], dim=0))


def megatron_qkv_reshape(param: torch.Tensor, head_size: int, n_heads: int) -> torch.Tensor:
    """Per-head interleaved ``[heads, 3, d]`` -> ``[q | k | v]``."""
    x = param.reshape(n_heads, 3, head_size, *param.shape[1:])
    return torch.cat([x[:, i].reshape(n_heads * head_size, *param.shape[1:]) for i in range(3)], dim=0)


class MegatronQKVParameter(ParameterBase):
    params: torch.Tensor

    def finalize(self) -> torch.Tensor:
        m = self.inference_model
        return m.transform_qkv_param(megatron_qkv_reshape(self.params, m.head_size, m.n_heads))


def transform_gqa_megatron(param: torch.Tensor, head_size: int, n_heads_q: int, n_heads_kv: int) -> torch.Tensor:
    """Grouped layout ``[kv_heads, (q_per_kv + 2), d]`` (Falcon-40B style) -> ``[q | k | v]``."""
    per = n_heads_q // n_heads_kv
    x = param.reshape(n_heads_kv, per + 2, head_size, *param.shape[1:])
    q = x[:, :per].reshape(n_heads_q * head_size, *param.shape[1:])
    k = x[:, per].reshape(n_heads_kv * head_size, *param.shape[1:])
    v = x[:, per + 1].reshape(n_heads_kv * head_size, *param.shape[1:])
    return torch.cat([q, k, v], dim=0)


class GQAMegatronQKVParameter(ParameterBase):
    params: torch.Tensor

    def finalize(self) -> torch.Tensor:
        m = self.inference_model
        return m.transform_qkv_param(transform_gqa_megatron(self.params, m.head_size, m.n_heads_q, m.n_heads_kv))
