"""Vocab cross-entropy over a sequence-sharded batch (reference ``sequence/cross_entropy.py:59``): every SP
rank computes the loss of its sequence shard; the per-token losses are all-gathered along the sequence so each
rank returns the full-sequence loss tensor, and the backward keeps only the local shard's gradient."""
import torch

from deepspeed_b200 import comm as dist
from deepspeed_b200.ops.kernels.transformer_ops import softmax_xent_fwd_bwd


class _VocabSequenceParallelCrossEntropy(torch.autograd.Function):

    @staticmethod
    def forward(ctx, vocab_seq_parallel_logits, target, sp_group):
        # logits: [S/P, B, V]  target: [S/P, B]
        sl, b, v = vocab_seq_parallel_logits.shape
        work = vocab_seq_parallel_logits.reshape(sl * b, v).clone()
        loss, grad = softmax_xent_fwd_bwd(work, target.reshape(-1).contiguous(), 1.0, None, -100, True)
        ctx.save_for_backward(grad)
        ctx.shape = (sl, b, v)
        ctx.sp_world = dist.get_world_size(sp_group)
        ctx.sp_rank = dist.get_rank(sp_group)
        loss = loss.view(sl, b)
        full = torch.empty(sl * ctx.sp_world, b, dtype=loss.dtype, device=loss.device)
        dist.all_gather_into_tensor(full, loss.contiguous(), group=sp_group)
        return full

    @staticmethod
    def backward(ctx, grad_output):
        (grad, ) = ctx.saved_tensors
        sl, b, v = ctx.shape
        go = grad_output[ctx.sp_rank * sl:(ctx.sp_rank + 1) * sl].reshape(sl * b, 1)
        return (grad.float() * go).to(grad.dtype).view(sl, b, v), None, None


def vocab_sequence_parallel_cross_entropy(vocab_parallel_logits, target, sp_group):
    return _VocabSequenceParallelCrossEntropy.apply(vocab_parallel_logits, target, sp_group)
