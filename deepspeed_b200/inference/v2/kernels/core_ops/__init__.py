"""Dense (non-ragged) kernels of the inference engine (reference ``inference/v2/kernels/core_ops``)."""
from .bias_activations.bias_activation import CUDABiasActivation  # noqa: F401
from .blas_kernels.blas_linear import BlasLibLinear  # noqa: F401
from .cuda_layer_norm.cuda_ln import CUDAFPLN  # noqa: F401
from .cuda_layer_norm.cuda_post_ln import CUDAFPPostLN  # noqa: F401
from .cuda_layer_norm.cuda_pre_ln import CUDAFPPreLN  # noqa: F401
from .cuda_linear.cuda_linear import CUDAWf6Af16Linear  # noqa: F401
from .cuda_rms_norm.rms_norm import CUDARMSNorm  # noqa: F401
from .cuda_rms_norm.rms_pre_norm import CUDARMSPreNorm  # noqa: F401
from .gated_activations.gated_activation import CUDAGatedActivation  # noqa: F401
