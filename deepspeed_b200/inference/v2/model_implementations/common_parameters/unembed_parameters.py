"""Reference ``model_implementations/common_parameters/unembed_parameters.py``."""
import torch

from ..parameter_base import ParameterBase, ParamList  # noqa: F401


class UnembedParameter(ParameterBase):
    """LM head [vocab, model_dim]."""
    params: torch.Tensor

    def finalize(self) -> torch.Tensor:
        return self.inference_model.transform_unembed_param(self.params)
