"""Reference ``model_implementations/common_parameters/embedding_parameters.py``."""
import torch

from ..parameter_base import ParameterBase, ParamList  # noqa: F401


class EmbeddingParameter(ParameterBase):
    """Token embedding [vocab, model_dim]."""
    params: torch.Tensor

    def finalize(self) -> torch.Tensor:
        return self.inference_model.transform_embedding_param(self.params)
