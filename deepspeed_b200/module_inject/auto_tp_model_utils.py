"""Model-level fix-ups AutoTP needs for families whose forward builds per-head tensors outside the attention module
(reference ``module_inject/auto_tp_model_utils.py``): ALiBi tensors must be built for the LOCAL heads only."""
import math

import torch

from deepspeed_b200 import comm as dist


def _alibi_slopes(n_heads, device):
    from deepspeed_b200.ops.transformer.inference.ds_transformer import alibi_slopes
    return alibi_slopes(n_heads, device)


def build_bloom_alibi_tensor(attention_mask: torch.Tensor, num_heads: int, dtype: torch.dtype) -> torch.Tensor:
    """BLOOM ALiBi bias ``[batch * local_heads, 1, seq]`` for this tensor-parallel rank's heads."""
    batch, seq = attention_mask.shape
    slopes = _alibi_slopes(num_heads, attention_mask.device)
    pos = ((attention_mask.cumsum(dim=-1) - 1) * attention_mask)[:, None, :].to(torch.float32)
    alibi = slopes[None, :, None] * pos  # [batch, heads, seq]
    if dist.is_initialized():
        from deepspeed_b200.utils import groups
        tp = groups._get_model_parallel_world_size()
        if tp > 1:
            per = num_heads // tp
            r = groups._get_model_parallel_rank()
            alibi = alibi[:, r * per:(r + 1) * per]
            num_heads = per
    return alibi.reshape(batch * num_heads, 1, seq).to(dtype)


def get_alibi_mask(self, tensor, seq_length_with_past):
    """Baichuan-style ALiBi mask (``[local_heads, seq, seq]`` additive, causal) for the local heads."""
    n = self.n_head
    slopes = _alibi_slopes(n, tensor.device)
    pos = torch.arange(seq_length_with_past, device=tensor.device, dtype=torch.float32)
    rel = pos[None, :] - pos[:, None]
    mask = slopes[:, None, None] * rel[None]
    mask = mask.masked_fill(rel[None] > 0, float("-inf"))
    if dist.is_initialized():
        from deepspeed_b200.utils import groups
        tp = groups._get_model_parallel_world_size()
        if tp > 1:
            per = n // tp
            r = groups._get_model_parallel_rank()
            mask = mask[r * per:(r + 1) * per]
    return mask.to(tensor.dtype)


def build_mpt_atten_bias_tensor(self, device, dtype, attention_mask=None, prefix_mask=None, sequence_id=None):
    """MPT builds its attention bias for all heads; keep this rank's heads."""
    attn_bias, attention_mask = self._attn_bias_orig(device=device, dtype=dtype, attention_mask=attention_mask,
                                                     prefix_mask=prefix_mask, sequence_id=sequence_id)
    if attn_bias is not None and dist.is_initialized():
        from deepspeed_b200.utils import groups
        tp = groups._get_model_parallel_world_size()
        if tp > 1:
            per = attn_bias.shape[1] // tp
            r = groups._get_model_parallel_rank()
            attn_bias = attn_bias[:, r * per:(r + 1) * per]
    return attn_bias, attention_mask


def build_mpt_alibi_tensor(self, num_heads, sequence_length, alibi_bias_max=8, device=None):
    alibi = self.build_mpt_alibi_tensor_orig(num_heads, sequence_length, alibi_bias_max, device)
    if dist.is_initialized():
        from deepspeed_b200.utils import groups
        tp = groups._get_model_parallel_world_size()
        if tp > 1:
            per = int(math.ceil(num_heads / tp))
            r = groups._get_model_parallel_rank()
            alibi = alibi[r * per:(r + 1) * per]
    return alibi
