"""Reference ``model_implementations/common_parameters/attn_output_parameters.py``."""
import torch

from ..parameter_base import ParameterBase, ParamList  # noqa: F401


class AttentionOutputParameter(ParameterBase):
    """Attention output projection [model_dim, heads * head_size] (row-parallel)."""
    params: torch.Tensor

    def finalize(self) -> torch.Tensor:
        return self.inference_model.transform_attn_out_param(self.params)
