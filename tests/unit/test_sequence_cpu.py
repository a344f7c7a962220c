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


def _dense_reference(x, w1, b1, w2, heads, kv_heads, d, rope=None):
    """Plain causal GQA attention over the full sequence ([S, B, H] in / out)."""
    S, B, _ = x.shape
    qkv = torch.nn.functional.linear(x, w1, b1)
    q, k, v = torch.split(qkv, [heads * d, kv_heads * d, kv_heads * d], dim=-1)
    q, k, v = q.reshape(S, B, heads, d), k.reshape(S, B, kv_heads, d), v.reshape(S, B, kv_heads, d)
    if rope is not None:
        from deepspeed_b200.sequence.layer import apply_rotary_pos_emb
        q, k = apply_rotary_pos_emb(q, *rope), apply_rotary_pos_emb(k, *rope)
    rep = heads // kv_heads
    k, v = k.repeat_interleave(rep, dim=2), v.repeat_interleave(rep, dim=2)
    q, k, v = (t.permute(1, 2, 0, 3) for t in (q, k, v))
    o = torch.nn.functional.scaled_dot_product_attention(q, k, v, is_causal=True).permute(2, 0, 1, 3).reshape(S, B, heads * d)
    return torch.nn.functional.linear(o, w2)


@pytest.mark.parametrize("kv_heads,use_rope", [(4, False), (2, True)])
def test_fpdt_manual_backward_matches_dense(kv_heads, use_rope):
    """The chunked core's hand-written backward (per-pair flash backward with the global LSE) against autograd through dense
    attention: input, QKV weight and bias gradients, with GQA and rotary embeddings."""
    from deepspeed_b200.sequence.fpdt_layer import FPDT_Attention
    torch.manual_seed(1)
    S, B, heads, d = 24, 2, 4, 8
    Hd = heads * d
    w1 = (torch.randn((heads + 2 * kv_heads) * d, Hd) * 0.2).requires_grad_(True)
    b1 = (torch.randn((heads + 2 * kv_heads) * d) * 0.1).requires_grad_(True)
    w2 = torch.randn(Hd, Hd) * 0.2
    rope = None
    if use_rope:
        ang = torch.arange(S)[:, None].float() * (1.0 / 10000**(torch.arange(0, d, 2).float() / d))[None]
        ang = torch.cat([ang, ang], -1)[:, None, None, :]
        rope = (ang.cos(), ang.sin())
    attn = FPDT_Attention(first_weight=w1, first_bias=b1, second_weight=w2, chunk_size=6, enable_offloading=False,
                          num_heads=heads, num_kv_heads=kv_heads, head_dim=d, return_bias=False)
    x = torch.randn(S, B, Hd, requires_grad=True)
    g = torch.randn(S, B, Hd)
    y = attn(x, rotary_pos_emb=rope)
    y.backward(g)
    got = [x.grad.clone(), w1.grad.clone(), b1.grad.clone()]
    x.grad = w1.grad = b1.grad = None
    ref = _dense_reference(x, w1, b1, w2, heads, kv_heads, d, rope)
    ref.backward(g)
    torch.testing.assert_close(y, ref, atol=2e-5, rtol=1e-4)
    for a, b_ in zip(got, (x.grad, w1.grad, b1.grad)):
        torch.testing.assert_close(a, b_, atol=5e-5, rtol=1e-4)


def _fpdt_sp_worker():
    """Two sequence-parallel ranks with the load-balanced FPDT chunk assignment reproduce single-process dense attention."""
    import torch.distributed as td
    from deepspeed_b200.sequence.fpdt_layer import FPDT_Attention, FPDT_InputConstruct
    torch.manual_seed(3)
    w, r = td.get_world_size(), td.get_rank()
    S, B, heads, d, n_chunks = 32, 1, 4, 8, 2
    Hd = heads * d
    w1 = (torch.randn(3 * Hd, Hd) * 0.2).requires_grad_(True)
    w2 = torch.randn(Hd, Hd) * 0.2
    full = torch.randn(S, B, Hd)
    g_full = torch.randn(S, B, Hd)
    ids = torch.arange(S)[None]  # [1, S]: which global positions this rank owns
    mine, *_ = FPDT_InputConstruct(ids, None, None, None, None, sp_size=w, sp_rank=r, num_chunks=n_chunks)
    mine = mine[0]
    x = full[mine].clone().requires_grad_(True)
    attn = FPDT_Attention(first_weight=w1, second_weight=w2, chunk_size=S // n_chunks, enable_offloading=False,
                          num_heads=heads, num_kv_heads=heads, head_dim=d, return_bias=False,
                          sequence_process_group=td.group.WORLD)
    y = attn(x)
    y.backward(g_full[mine])
    xf = full.clone().requires_grad_(True)
    w1f = w1.detach().clone().requires_grad_(True)
    ref = _dense_reference(xf, w1f, None, w2, heads, heads, d)
    ref.backward(g_full)
    torch.testing.assert_close(y, ref[mine], atol=2e-5, rtol=1e-4)
    torch.testing.assert_close(x.grad, xf.grad[mine], atol=5e-5, rtol=1e-4)
    wg = w1.grad.clone()
    td.all_reduce(wg)  # the weight gradient is a sum over the sequence shards
    torch.testing.assert_close(wg, w1f.grad, atol=1e-4, rtol=1e-4)


def test_fpdt_two_rank_sequence_parallel_matches_dense():
    run_distributed(_fpdt_sp_worker, 2)


# ---- engine-level Ulysses: sequence_parallel_size in the config, ZeRO over the seq x data group ------------------------------
from torch import nn  # noqa: E402


class SPBlock(nn.Module):
    """One attention + MLP block whose attention runs sequence-parallel (Ulysses)."""

    def __init__(self, d, heads, sp):
        super().__init__()
        self.qkv = nn.Linear(d, 3 * d)
        self.out = nn.Linear(d, d)
        self.mlp = nn.Linear(d, d)
        self.heads = heads
        self.sp = sp
        if sp:
            from deepspeed_b200.sequence.layer import DistributedAttention
            self.attn = DistributedAttention(self._local, None, scatter_idx=2, gather_idx=0)

    @staticmethod
    def _local(q, k, v):  # [S, B, h, D]
        qq, kk, vv = (t.permute(1, 2, 0, 3) for t in (q, k, v))
        return torch.nn.functional.scaled_dot_product_attention(qq, kk, vv, is_causal=True).permute(2, 0, 1, 3)

    def forward(self, x, y):  # x [S_local, B, d]
        S, B, d = x.shape
        q, k, v = (t.reshape(S, B, self.heads, d // self.heads) for t in self.qkv(x).chunk(3, -1))
        a = self.attn(q, k, v) if self.sp else self._local(q, k, v)
        h = x + self.out(a.reshape(S, B, d))
        h = h + torch.tanh(self.mlp(h))
        return ((h - y)**2).mean()


def _ulysses_engine_worker():
    import deepspeed_b200 as ds
    from deepspeed_b200.utils import safe_get_full_fp32_param
    w, r = ds.comm.get_world_size(), ds.comm.get_rank()
    torch.manual_seed(0)
    d, heads, S, B = 16, 4, 8 * w, 2
    ref = SPBlock(d, heads, sp=False)
    model = SPBlock(d, heads, sp=True)
    model.load_state_dict(ref.state_dict())
    cfg = {"train_micro_batch_size_per_gpu": B, "sequence_parallel_size": w,
           "optimizer": {"type": "SGD", "params": {"lr": 0.1}}, "zero_optimization": {"stage": 1}}
    eng, *_ = ds.initialize(model=model, config=cfg)
    assert eng.sequence_parallel_size == w and eng.seq_dp_world_size == w and eng.dp_world_size == 1
    ropt = torch.optim.SGD(ref.parameters(), lr=0.1)
    g = torch.Generator().manual_seed(3)
    sl = slice(r * S // w, (r + 1) * S // w)
    for it in range(4):
        x, y = torch.randn(S, B, d, generator=g), torch.randn(S, B, d, generator=g)
        loss = eng(x[sl], y[sl])   # this rank's sequence shard; the loss is the mean over the LOCAL tokens
        eng.backward(loss)
        eng.step()
        ref(x, y).backward()
        ropt.step(); ropt.zero_grad()
    for (n, p), (_, q) in zip(model.named_parameters(), ref.named_parameters()):
        torch.testing.assert_close(safe_get_full_fp32_param(p).cpu(), q.detach(), atol=2e-5, rtol=1e-4, msg=n)



def test_engine_sequence_parallel_matches_full_sequence_training():
    """``sequence_parallel_size`` = world: every rank trains on its sequence shard through ``DistributedAttention``; ZeRO-1
    shards and reduces over the sequence x data group (reference ``engine.py:1655``); parameters track single-process
    training on the full sequence (SGD: the key bias gradient is analytically zero, Adam would amplify its rounding noise)."""
    run_distributed(_ulysses_engine_worker, 2)
