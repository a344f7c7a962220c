"""Reference ``model_implementations/common_parameters/norm_parameters.py``."""
import torch

from ..parameter_base import ParameterBase, ParamList  # noqa: F401


class NormParameter(ParameterBase):
    """LayerNorm / RMSNorm gamma or beta."""
    params: torch.Tensor

    def finalize(self) -> torch.Tensor:
        return self.inference_model.transform_norm_param(self.params)
