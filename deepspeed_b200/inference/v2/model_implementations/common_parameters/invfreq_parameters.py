"""Reference ``model_implementations/common_parameters/invfreq_parameters.py``."""
import torch

from ..parameter_base import ParameterBase, ParamList  # noqa: F401


class InvFreqParameter(ParameterBase):
    """Trained rotary inverse frequencies."""
    params: torch.Tensor

    def finalize(self) -> torch.Tensor:
        return self.params.to(self.inference_model.activation_dtype)
