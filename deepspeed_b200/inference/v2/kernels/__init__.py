"""Kernel namespace of the ragged engine (reference ``inference/v2/kernels/{core_ops,ragged_ops,cutlass_ops}``).
All device code lives in ``libdsb200_cuda.so``; this module re-exports the Python faces under the names the
reference's module layer uses."""
from deepspeed_b200.ops.kernels.ragged_ops import (kv_rotary_append as linear_blocked_kv_rotary,  # noqa: F401
                                                   paged_attention as blocked_flash, ragged_embed, row_gather as
                                                   logits_gather)
from deepspeed_b200.ops.kernels.moe_ops import top_k_gating, scatter as moe_scatter, gather as moe_gather  # noqa: F401
from deepspeed_b200.ops.kernels.transformer_ops import (rms_norm, layer_norm, gated_act as gated_activation,  # noqa: F401
                                                        bias_act as bias_activation)
from deepspeed_b200.ops.gemm import matmul_nt as blas_linear  # noqa: F401
