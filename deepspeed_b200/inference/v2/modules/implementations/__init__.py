"""Importing this package registers every implementation with its registry."""
from .attention.dense_blocked_attention import DSDenseBlockedAttention  # noqa: F401
from .embedding.ragged_embedding import DSRaggedEmbedding  # noqa: F401
from .linear.blas_fp_linear import BlasFPLinear  # noqa: F401
from .linear.quantized_linear import QuantizedWf6Af16Linear  # noqa: F401
from .moe.cutlass_multi_gemm import DSMultiGemmMoE  # noqa: F401
from .post_norm.cuda_post_ln import DSPostLNCUDAModule  # noqa: F401
from .pre_norm.cuda_pre_ln import DSPreLNCUDAModule  # noqa: F401
from .pre_norm.cuda_pre_rms import DSPreRMSCUDAModule  # noqa: F401
from .unembed.ragged_unembed import DSRaggedUnembed  # noqa: F401
