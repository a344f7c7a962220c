"""Ring / context-parallel attention (absent from the reference, SURVEY.md 5.7 -- added natively).

Each rank owns a contiguous sequence block of Q, K, V.  K/V blocks travel around the ring (P-1 hops of
``batch_isend_irecv``); after every hop the local Q attends to the visiting block and the partial results are
merged with the log-sum-exp rule (same online-softmax algebra as ``fpdt_layer.update_out_and_lse``).  Causal
masking skips blocks from the future and applies the triangular mask only on the diagonal block.  On an
NVSwitch box every hop is a full-bandwidth peer copy, so the ring is bandwidth-equivalent to an all-gather of
K/V but keeps the memory footprint at two blocks.
"""
import math

import torch

from deepspeed_b200 import comm as dist


def _block_attn(q, k, v, causal_diag, scale):
    """Returns (out [B,H,S,D] fp32, lse [B,H,S] fp32) for one K/V block."""
    s = torch.matmul(q.float(), k.float().transpose(-1, -2)) * scale
    if causal_diag:
        Sq, Sk = s.shape[-2:]
        mask = torch.ones(Sq, Sk, dtype=torch.bool, device=s.device).tril_()
        s = s.masked_fill(~mask, float("-inf"))
    lse = torch.logsumexp(s, dim=-1)
    p = torch.exp(s - lse[..., None])
    return torch.matmul(p, v.float()), lse


def _merge(out, lse, o2, l2):
    if out is None:
        return o2, l2
    new = torch.logaddexp(lse, l2)
    out = out * torch.exp(lse - new)[..., None] + o2 * torch.exp(l2 - new)[..., None]
    return out, new


class _RingAttention(torch.autograd.Function):

    @staticmethod
    def forward(ctx, q, k, v, group, causal, scale):
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        scale = scale or 1.0 / math.sqrt(q.shape[-1])
        out = lse = None
        kb, vb = k.contiguous(), v.contiguous()
        src = rank
        for hop in range(world):
            if not causal or src <= rank:
                o2, l2 = _block_attn(q, kb, vb, causal and src == rank, scale)
                out, lse = _merge(out, lse, o2, l2)
            if hop < world - 1:
                kb, vb = _rotate(kb, vb, group, rank, world)
                src = (src - 1) % world
        ctx.save_for_backward(q, k, v, out.to(q.dtype), lse)
        ctx.group, ctx.causal, ctx.scale = group, causal, scale
        return out.to(q.dtype)

    @staticmethod
    def backward(ctx, dout):
        q, k, v, out, lse = ctx.saved_tensors
        group, causal, scale = ctx.group, ctx.causal, ctx.scale
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        delta = (dout.float() * out.float()).sum(-1)  # [B,H,S]
        dq = torch.zeros_like(q, dtype=torch.float32)
        kb, vb = k.contiguous(), v.contiguous()
        dkb = torch.zeros_like(k, dtype=torch.float32)
        dvb = torch.zeros_like(v, dtype=torch.float32)
        src = rank
        for hop in range(world):
            if not causal or src <= rank:
                s = torch.matmul(q.float(), kb.float().transpose(-1, -2)) * scale
                if causal and src == rank:
                    Sq, Sk = s.shape[-2:]
                    s = s.masked_fill(~torch.ones(Sq, Sk, dtype=torch.bool, device=s.device).tril_(), float("-inf"))
                p = torch.exp(s - lse[..., None])
                dvb += torch.matmul(p.transpose(-1, -2), dout.float())
                dp = torch.matmul(dout.float(), vb.float().transpose(-1, -2))
                ds = p * (dp - delta[..., None]) * scale
                dq += torch.matmul(ds, kb.float())
                dkb += torch.matmul(ds.transpose(-1, -2), q.float())
            # the K/V block travels on together with its gradient accumulators; after P hops both are home
            kb, vb, dkb, dvb = _rotate4(kb, vb, dkb, dvb, group, rank, world)
            src = (src - 1) % world
        return dq.to(q.dtype), dkb.to(k.dtype), dvb.to(v.dtype), None, None, None


def _rotate(a, b, group, rank, world):
    nxt = dist.get_global_rank(group, (rank + 1) % world)
    prv = dist.get_global_rank(group, (rank - 1) % world)
    ra, rb = torch.empty_like(a), torch.empty_like(b)
    ops = [dist.P2POp(torch.distributed.isend, a, nxt, group), dist.P2POp(torch.distributed.isend, b, nxt, group),
           dist.P2POp(torch.distributed.irecv, ra, prv, group), dist.P2POp(torch.distributed.irecv, rb, prv, group)]
    for w in dist.batch_isend_irecv(ops):
        w.wait()
    return ra, rb


def _rotate4(a, b, c, d, group, rank, world):
    a2, b2 = _rotate(a, b, group, rank, world)
    c2, d2 = _rotate(c.contiguous(), d.contiguous(), group, rank, world)
    return a2, b2, c2, d2


def ring_attention(q, k, v, group=None, causal=True, scale=None):
    """q, k, v: ``[B, H, S_local, D]`` (this rank's contiguous sequence block)."""
    if dist.get_world_size(group) == 1:
        return torch.nn.functional.scaled_dot_product_attention(q, k, v, is_causal=causal, scale=scale)
    return _RingAttention.apply(q, k, v, group, causal, scale)
