"""LM head over a ragged batch (reference ``modules/implementations/unembed/ragged_unembed.py``): gather each sequence's
last token, final norm, vocabulary projection."""
from typing import Any, Dict

import torch

from ....inference_utils import DtypeEnum, NormTypeEnum
from ....kernels.core_ops import BlasLibLinear, CUDAFPLN, CUDARMSNorm
from ....kernels.ragged_ops import RaggedLogitsGather
from ...configs import DSUnembedConfig
from ...interfaces import DSUnembedBase, DSUnembedRegistry


@DSUnembedRegistry.register_module
class DSRaggedUnembed(DSUnembedBase):

    @staticmethod
    def name() -> str:
        return "ragged_unembed"

    @staticmethod
    def supports_config(config: DSUnembedConfig) -> bool:
        return True

    def __init__(self, config: DSUnembedConfig, implementation_config: Dict[str, Any] = None) -> None:
        super().__init__(config, implementation_config)
        dt = DtypeEnum(config.dtype).value
        self.gather = RaggedLogitsGather(config.model_dim, dt)
        self.norm = None
        if config.norm_type is not None:
            self.norm = (CUDARMSNorm if NormTypeEnum(config.norm_type) == NormTypeEnum.RMSNorm else CUDAFPLN)(config.model_dim, dt)
        self.gemm = BlasLibLinear(dt)
        self._out = None

    @property
    def output(self) -> torch.Tensor:
        return self._out

    def forward(self, hidden_states, vocab_embedding, ragged_metadata, bias=None, gamma=None, beta=None) -> torch.Tensor:
        idx = ragged_metadata.last_token_index()
        last = self.gather(torch.empty(idx.numel(), hidden_states.shape[1], dtype=hidden_states.dtype,
                                       device=hidden_states.device), hidden_states, idx)
        if self.norm is not None:
            normed = torch.empty_like(last)
            last = self.norm(normed, last, gamma) if isinstance(self.norm, CUDARMSNorm) else self.norm(normed, last, gamma, beta)
        logits = torch.empty(last.shape[0], vocab_embedding.shape[0], dtype=last.dtype, device=last.device)
        self.gemm(logits, last, vocab_embedding)
        self._out = logits if bias is None else logits + bias
        return self._out
