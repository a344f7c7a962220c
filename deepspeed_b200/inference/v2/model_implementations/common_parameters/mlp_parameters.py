"""Reference ``model_implementations/common_parameters/mlp_parameters.py``."""
import torch

from ..parameter_base import ParameterBase, ParamList  # noqa: F401


class MLP1Parameter(ParameterBase):
    """First MLP projection (no gating)."""
    params: torch.Tensor

    def finalize(self) -> torch.Tensor:
        return self.inference_model.transform_mlp_1_param(self.params)


class GatedMLPParameter(ParameterBase):
    """Gated first projection: ``gate`` and ``up`` are fused by stacking [gate; up]."""
    gate_params: torch.Tensor
    up_params: torch.Tensor

    def finalize(self) -> torch.Tensor:
        return self.inference_model.transform_mlp_1_param(torch.cat([self.gate_params, self.up_params], dim=0))


class MLP2Parameter(ParameterBase):
    """Second MLP projection (row-parallel)."""
    params: torch.Tensor

    def finalize(self) -> torch.Tensor:
        return self.inference_model.transform_mlp_2_param(self.params)


class FusedGatedMLPParameter(ParameterBase):
    """Checkpoints that already store ``[gate; up]`` as one matrix (Phi-3 ``gate_up_proj``): split, then hand both halves
    to the model in the order its fused SwiGLU kernel expects (reference ``mlp_parameters.py:69``)."""
    params: torch.Tensor

    def finalize(self) -> torch.Tensor:
        half = self.params.shape[0] // 2
        return self.inference_model.transform_mlp_1_param(torch.cat([self.params[:half], self.params[half:]], dim=0))
