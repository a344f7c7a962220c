"""Gather the hidden state of the LAST token of every sequence in the ragged batch (the only rows the unembedding needs).

Reference ``inference/v2/kernels/ragged_ops/logits_gather/logits_gather.py``."""
import torch

from deepspeed_b200.ops.kernels import ragged_ops as R

from ...ds_kernel import DSKernelBase, check_dtype


class RaggedLogitsGather(DSKernelBase):

    def __init__(self, model_dim: int, fp_dtype):
        check_dtype(fp_dtype, "RaggedLogitsGather")
        if model_dim * torch.empty(0, dtype=fp_dtype).element_size() % 16 != 0:
            raise ValueError("model_dim must be a multiple of 16 bytes")

    def __call__(self, final_token_activations, all_activations, last_token_index) -> torch.Tensor:
        final_token_activations.copy_(R.row_gather(all_activations, last_token_index))
        return final_token_activations
