"""Sequence parallelism on the host tier: Ulysses all-to-all round trips (even / uneven heads), distributed
attention vs single-device attention, ring attention fwd/bwd, FPDT chunked attention vs dense attention."""
import math

import pytest
import torch

from tests.common import run_distributed


def _ulysses_roundtrip(heads):
    import torch.distributed as td
    from deepspeed_b200.sequence.layer import DistributedAttention, _SeqAllToAll
    torch.manual_seed(0)
    w, r = td.get_world_size(), td.get_rank()
    S, B, H, D = 8 * w, 2, heads, 4
    full_q, full_k, full_v = (torch.randn(S, B, H, D) for _ in range(3))
    sl = slice(r * S // w, (r + 1) * S // w)
    q, k, v = (t[sl].clone().requires_grad_(True) for t in (full_q, full_k, full_v))
    x = _SeqAllToAll.apply(None, q, 2, 0, 1)
    back = _SeqAllToAll.apply(None, x, 0, 2, 1)
    assert torch.equal(back, q)

    def local_attn(q_, k_, v_):
        # [S, B, h, D] -> attention over S
        qq, kk, vv = (t.permute(1, 2, 0, 3) for t in (q_, k_, v_))
        o = torch.nn.functional.scaled_dot_product_attention(qq, kk, vv, is_causal=True)
        return o.permute(2, 0, 1, 3)

    out = DistributedAttention(local_attn, None, scatter_idx=2, gather_idx=0)(q, k, v)
    ref = local_attn(full_q, full_k, full_v)[sl]
    assert (out - ref).abs().max() < 1e-5
    out.sum().backward()
    fq = full_q.clone().requires_grad_(True)
    local_attn(fq, full_k, full_v).sum().backward()
    assert (q.grad - fq.grad[sl]).abs().max() < 1e-5


@pytest.mark.parametrize("heads", [4, 3])
def test_ulysses_even_and_uneven_heads(heads):
    run_distributed(_ulysses_roundtrip, 2, (heads, ))


def _ring():
    import torch.distributed as td
    from deepspeed_b200.sequence.ring_attention import ring_attention
    torch.manual_seed(0)
    w, r = td.get_world_size(), td.get_rank()
    B, H, S, D = 2, 3, 6 * w, 8
    fq, fk, fv = (torch.randn(B, H, S, D) for _ in range(3))
    sl = slice(r * S // w, (r + 1) * S // w)
    q, k, v = (t[:, :, sl].clone().requires_grad_(True) for t in (fq, fk, fv))
    out = ring_attention(q, k, v, None, causal=True)
    rq, rk, rv = (t.clone().requires_grad_(True) for t in (fq, fk, fv))
    ref = torch.nn.functional.scaled_dot_product_attention(rq, rk, rv, is_causal=True)
    assert (out - ref[:, :, sl]).abs().max() < 1e-5
    g = torch.randn_like(ref)
    ref.backward(g)
    out.backward(g[:, :, sl])
    assert (q.grad - rq.grad[:, :, sl]).abs().max() < 1e-4
    assert (k.grad - rk.grad[:, :, sl]).abs().max() < 1e-4
    assert (v.grad - rv.grad[:, :, sl]).abs().max() < 1e-4


def test_ring_attention_matches_dense():
    run_distributed(_ring, 2)


def test_fpdt_chunked_attention_matches_dense():
    from deepspeed_b200.sequence.fpdt_layer import FPDT_Attention, FPDT_FFN, FPDT_LogitsLoss, update_out_and_lse
    torch.manual_seed(0)
    S, B, Hd, heads = 32, 2, 16, 4
    w1 = torch.randn(3 * Hd, Hd) * 0.2
    w2 = torch.randn(Hd, Hd) * 0.2
    attn = FPDT_Attention(first_weight=w1, second_weight=w2, chunk_size=8, enable_offloading=False, num_heads=heads,
                          num_kv_heads=heads, head_dim=Hd // heads, return_bias=False)
    x = torch.randn(S, B, Hd, requires_grad=True)
    y = attn(x)
    qkv = torch.nn.functional.linear(x, w1)
    q, k, v = (t.reshape(S, B, heads, Hd // heads).permute(1, 2, 0, 3) for t in qkv.chunk(3, dim=-1))
    ref = torch.nn.functional.scaled_dot_product_attention(q, k, v, is_causal=True).permute(2, 0, 1, 3).reshape(S, B, Hd)
    ref = torch.nn.functional.linear(ref, w2)
    assert (y - ref).abs().max() < 1e-5
    y.sum().backward()
    assert x.grad is not None and torch.isfinite(x.grad).all()
    ffn = FPDT_FFN(torch.randn(4 * Hd, Hd) * 0.1, torch.zeros(4 * Hd), torch.randn(Hd, 4 * Hd) * 0.1, torch.zeros(Hd),
                   chunk_size=8)
    z = ffn(x.detach())
    assert z.shape == (S, B, Hd)
    lw = torch.randn(50, Hd)
    labels = torch.randint(0, 50, (S, B))
    loss = FPDT_LogitsLoss(lw, chunk_size=16)(x.detach(), labels)
    ref_loss = torch.nn.functional.cross_entropy((x.detach().reshape(-1, Hd) @ lw.t()), labels.reshape(-1))
    assert abs(loss.item() - ref_loss.item()) < 1e-5
