"""One import point for every parallelism strategy.

=====================  =========================================================================================
data parallel / ZeRO   ``deepspeed_b200.runtime.zero`` (``ZeroShardedOptimizer`` stages 0-3, MiCS, ZeRO++)
tensor parallel        ``AutoTP`` / ``tp_model_init`` / ``LinearLayer`` / ``LinearAllreduce``; ``DominoTransformerLayer``
pipeline parallel      ``PipelineModule`` / ``LayerSpec`` / ``TiedLayerSpec`` / ``PipelineEngine`` / schedules
expert parallel        ``MoE`` layer, ``TopKGate``, expert groups
sequence parallel      ``DistributedAttention`` (Ulysses), ``ring_attention``, ``FPDT`` chunked offload
symmetric collectives  ``deepspeed_b200.comm.symm`` (NVLink peer / NVLS fused kernels)
=====================  =========================================================================================
"""
from deepspeed_b200.module_inject.auto_tp import AutoTP, tp_model_init  # noqa: F401
from deepspeed_b200.module_inject.layers import LinearAllreduce, LinearLayer, LmHeadLinearAllreduce  # noqa: F401
from deepspeed_b200.moe.layer import MoE  # noqa: F401
from deepspeed_b200.runtime.domino import DominoTransformer, DominoTransformerLayer  # noqa: F401
from deepspeed_b200.runtime.pipe import LayerSpec, PipelineModule, TiedLayerSpec  # noqa: F401
from deepspeed_b200.runtime.pipe.topology import PipeDataParallelTopology, PipeModelDataParallelTopology, ProcessTopology  # noqa: F401
from deepspeed_b200.runtime.zero.mics import MiCS_Init, MiCS_Optimizer  # noqa: F401
from deepspeed_b200.runtime.zero.sharded import ZeroShardedOptimizer  # noqa: F401
from deepspeed_b200.sequence.layer import DistributedAttention  # noqa: F401
from deepspeed_b200.utils import groups  # noqa: F401
