"""Top-k MoE block as gating -> scatter -> grouped GEMM -> activation -> grouped GEMM -> gather (reference
``modules/implementations/moe/cutlass_multi_gemm.py``)."""
from typing import Any, Dict

import torch

from deepspeed_b200.ops.kernels import transformer_ops as T

from ....inference_utils import ActivationType, DtypeEnum, is_gated
from ....kernels.cutlass_ops import MoEGEMM
from ....kernels.ragged_ops import MoEGather, MoEScatter, RaggedTopKGating
from ...configs import DSMoEConfig
from ...interfaces import DSMoEBase, DSMoERegistry
from ..linear.blas_fp_linear import _ACT


@DSMoERegistry.register_module
class DSMultiGemmMoE(DSMoEBase):

    @staticmethod
    def name() -> str:
        return "cutlass_multi_gemm_moe"

    @staticmethod
    def supports_config(config: DSMoEConfig) -> bool:
        return config.input_dtype == config.output_dtype and 1 <= config.top_k <= config.n_experts

    def __init__(self, config: DSMoEConfig, implementation_config: Dict[str, Any] = None) -> None:
        super().__init__(config, implementation_config)
        dt = DtypeEnum(config.input_dtype).value
        self.act = ActivationType(config.activation)
        self.gate = RaggedTopKGating(dt)
        self.scatter = MoEScatter(dt, config.model_dim)
        self.gemm = MoEGEMM(dt)
        self.gather = MoEGather(dt, config.model_dim, config.normalize_scores)
        self._out = None

    @property
    def output(self) -> torch.Tensor:
        return self._out

    def forward(self, hidden_states, gate_w, mlp_1_w, mlp_2_w, mlp_1_b=None, mlp_2_b=None) -> torch.Tensor:
        c = self._config
        T_, dev = hidden_states.shape[0], hidden_states.device
        k, E = c.top_k, c.n_experts
        logits = torch.nn.functional.linear(hidden_states, gate_w)
        counts = torch.zeros(E, dtype=torch.int32, device=dev)
        scores = torch.empty(T_, k, dtype=torch.float32, device=dev)
        assign = torch.empty(T_, k, dtype=torch.int32, device=dev)
        offs = torch.empty(T_, k, dtype=torch.int32, device=dev)
        self.gate(counts, scores, assign, offs, logits)
        moe_in = torch.empty(T_ * k, c.model_dim, dtype=hidden_states.dtype, device=dev)
        cum = torch.empty(E, dtype=torch.int32, device=dev)
        slots = torch.empty(T_, k, dtype=torch.int32, device=dev)
        self.scatter(moe_in, cum, slots, hidden_states, counts, assign, offs)
        inter = torch.empty(T_ * k, mlp_1_w.shape[1], dtype=hidden_states.dtype, device=dev)
        self.gemm(inter, moe_in, mlp_1_w, cum, mlp_1_b)
        inter = T.gated_act(inter.contiguous(), act=_ACT[self.act]) if is_gated(self.act) else (
            T.bias_act(inter, None, act=_ACT[self.act]) if _ACT[self.act] else inter)
        out_rows = torch.empty(T_ * k, c.model_dim, dtype=hidden_states.dtype, device=dev)
        self.gemm(out_rows, inter, mlp_2_w, cum, mlp_2_b)
        self._out = self.gather(torch.empty_like(hidden_states), out_rows, scores, slots)
        return self._out
