"""Random layerwise token dropping ops (reference ``ops/random_ltd/dropping_utils.py``, kernels N14)."""
import torch

from deepspeed_b200.ops.kernels import misc_ops as K


def gpt_sample_tokens(reserved_length: int, seq_length: int, batch_size: int, layers: int = 1, device="cpu",
                      attn_mask: torch.Tensor = None):
    """Per layer/batch: ``reserved_length`` distinct sorted token ids.  Returns (indices [L,B,k], new causal mask)."""
    prob = torch.ones(layers * batch_size, seq_length, device=device)
    idx = torch.multinomial(prob, reserved_length).to(torch.int32)
    idx = K.token_sort_(idx.contiguous()).reshape(layers, batch_size, reserved_length)
    new_mask = attn_mask[:, :, :reserved_length, :reserved_length] if attn_mask is not None else None
    return idx, new_mask


def bert_sample_tokens(reserved_length: int, seq_length: int, batch_size: int, layers: int = 1, device="cpu",
                       attn_mask: torch.Tensor = None):
    assert attn_mask is not None
    prob = torch.ones(layers * batch_size, seq_length, device=device)
    idx = torch.multinomial(prob, reserved_length).to(torch.int32)
    idx = K.token_sort_(idx.contiguous()).reshape(layers, batch_size, reserved_length)
    masks = [K.mask_gather(attn_mask.to(torch.float32) if not attn_mask.is_floating_point() else attn_mask, idx[l])
             .to(attn_mask.dtype) for l in range(layers)]
    return idx, masks


class GatherTokens(torch.autograd.Function):

    @staticmethod
    def forward(ctx, activations: torch.Tensor, sorted_indices: torch.Tensor, batch_first: bool):
        ctx.batch_first = batch_first
        x = activations if batch_first else activations.transpose(0, 1)
        ctx.shape = x.shape
        ctx.save_for_backward(sorted_indices)
        out = K.token_gather(x.contiguous(), sorted_indices)
        return activations, (out if batch_first else out.transpose(0, 1))

    @staticmethod
    def backward(ctx, a_grad, g_grad):
        (idx, ) = ctx.saved_tensors
        g = g_grad if ctx.batch_first else g_grad.transpose(0, 1)
        full = a_grad.clone() if ctx.batch_first else a_grad.transpose(0, 1).clone()
        B, k, H = g.shape
        # accumulate: the gathered tokens' gradient adds onto the pass-through gradient
        full.scatter_add_(1, idx.long()[..., None].expand(B, k, H), g.contiguous())
        return (full if ctx.batch_first else full.transpose(0, 1)), None, None


class ScatterTokens(torch.autograd.Function):

    @staticmethod
    def forward(ctx, all_activations: torch.Tensor, layer_activations: torch.Tensor, sorted_indices: torch.Tensor,
                batch_first: bool):
        ctx.batch_first = batch_first
        ctx.save_for_backward(sorted_indices)
        full = (all_activations if batch_first else all_activations.transpose(0, 1)).contiguous().clone()
        part = (layer_activations if batch_first else layer_activations.transpose(0, 1)).contiguous()
        K.token_scatter_(full, part, sorted_indices)
        return full if batch_first else full.transpose(0, 1)

    @staticmethod
    def backward(ctx, out_grad):
        (idx, ) = ctx.saved_tensors
        g = (out_grad if ctx.batch_first else out_grad.transpose(0, 1)).contiguous()
        part = K.token_gather(g, idx)
        full = g.clone()
        B, k, H = part.shape
        full.scatter_(1, idx.long()[..., None].expand(B, k, H), torch.zeros_like(part))
        if not ctx.batch_first:
            full, part = full.transpose(0, 1), part.transpose(0, 1)
        return full, part, None, None


def mask_gather_gpt(attn_mask, reserved_length):
    """Causal masks only need their top-left corner (reference ``mask_gather_gpt``)."""
    return attn_mask[:, :, :reserved_length, :reserved_length]


def mask_gather_bert(attn_mask, sorted_indices):
    """Slice a full [B,1,S,S] mask down to the kept tokens (reference ``mask_gather_bert``)."""
    return K.mask_gather(attn_mask, sorted_indices)
