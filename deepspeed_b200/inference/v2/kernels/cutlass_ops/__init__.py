"""Quantised / grouped GEMMs (reference ``inference/v2/kernels/cutlass_ops``)."""
from .mixed_gemm.mixed_gemm import MixedGEMM  # noqa: F401
from .moe_gemm.moe_gemm import MoEGEMM  # noqa: F401
from .moe_gemm.mixed_moe_gemm import MixedMoEGEMM  # noqa: F401
