"""Reference ``model_implementations/common_parameters/moe_parameters.py``."""
import torch

from ..parameter_base import ParameterBase, ParamList  # noqa: F401


class MoEGatingWeightParameter(ParameterBase):
    """Router [n_experts, model_dim]."""
    params: torch.Tensor

    def finalize(self) -> torch.Tensor:
        return self.inference_model.transform_moe_gate_param(self.params)


class UnfusedMoEMLP1Parameter(ParameterBase):
    """Per-expert first projections stacked into [n_experts, out, in]."""
    experts = ParamList("n_experts")

    def finalize(self) -> torch.Tensor:
        return self.inference_model.transform_moe_mlp_1_param(torch.stack(list(self.experts), dim=0))


class UnfusedMoEMLP2Parameter(ParameterBase):
    experts = ParamList("n_experts")

    def finalize(self) -> torch.Tensor:
        return self.inference_model.transform_moe_mlp_2_param(torch.stack(list(self.experts), dim=0))


class UnfusedMoEGatedMLPParameter(ParameterBase):
    """Per-expert gate / up pairs -> [n_experts, 2 * intermediate, model_dim]."""
    gating_experts = ParamList("n_experts")
    up_experts = ParamList("n_experts")

    def finalize(self) -> torch.Tensor:
        fused = torch.stack([torch.cat([g, u], dim=0) for g, u in zip(self.gating_experts, self.up_experts)], dim=0)
        return self.inference_model.transform_moe_mlp_1_param(fused)
