"""``deepspeed.ops.op_builder`` import path (a symlink to the top-level ``op_builder`` package in the reference tree): every
name and sub-module resolves to ``deepspeed_b200.op_builder``."""
import importlib
import sys

from deepspeed_b200.op_builder import *  # noqa: F401,F403
from deepspeed_b200.op_builder import ALL_OPS, OpBuilder, CUDAOpBuilder, CPUOpBuilder  # noqa: F401
from deepspeed_b200 import op_builder as _real

for _n in dir(_real):
    if _n.endswith("Builder"):
        globals()[_n] = getattr(_real, _n)
for _sub in ['all_ops', 'async_io', 'builder', 'cpu_adagrad', 'cpu_adam', 'cpu_lion', 'evoformer_attn', 'fp_quantizer', 'fused_adam', 'fused_lamb', 'fused_lion', 'gds', 'inference_core_ops', 'inference_cutlass_builder', 'quantizer', 'ragged_ops', 'ragged_utils', 'random_ltd', 'sparse_attn', 'spatial_inference', 'stochastic_transformer', 'transformer', 'transformer_inference']:
    sys.modules[f"{__name__}.{_sub}"] = importlib.import_module(f"deepspeed_b200.op_builder.{_sub}")
    globals()[_sub] = sys.modules[f"{__name__}.{_sub}"]
