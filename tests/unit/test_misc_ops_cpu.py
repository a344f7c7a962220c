"""Host definitions of the misc ops (the GPU tests compare the kernels against these), sparse attention,
evoformer attention, random-LTD, flatten, shared-memory collectives."""
import math

import pytest
import torch

from tests.common import run_distributed


def test_sparse_attention_matches_masked_dense():
    from deepspeed_b200.ops.sparse_attention import (BigBirdSparsityConfig, BSLongformerSparsityConfig, DenseSparsityConfig,
                                                     FixedSparsityConfig, LocalSlidingWindowSparsityConfig,
                                                     SparseSelfAttention, VariableSparsityConfig)
    torch.manual_seed(0)
    B, H, S, D, blk = 2, 4, 64, 8, 8
    q, k, v = (torch.randn(B, H, S, D, requires_grad=True) for _ in range(3))
    cfgs = [DenseSparsityConfig(H, blk), FixedSparsityConfig(H, blk, num_local_blocks=4, num_global_blocks=1),
            FixedSparsityConfig(H, blk, num_local_blocks=4, attention="unidirectional"),
            VariableSparsityConfig(H, blk, num_random_blocks=1, local_window_blocks=[2, 4], global_block_indices=[0]),
            BigBirdSparsityConfig(H, blk, num_random_blocks=1, num_sliding_window_blocks=3, num_global_blocks=1),
            BSLongformerSparsityConfig(H, blk, num_sliding_window_blocks=3, global_block_indices=[0, 5]),
            LocalSlidingWindowSparsityConfig(H, blk, num_sliding_window_blocks=3)]
    kpm = torch.zeros(B, S)
    kpm[:, -5:] = -10000.0
    for cfg in cfgs:
        att = SparseSelfAttention(cfg, max_seq_length=S)
        layout = att.get_layout(S)
        assert layout.shape == (H, S // blk, S // blk)
        if isinstance(cfg, FixedSparsityConfig) and cfg.attention == "unidirectional":
            assert torch.equal(layout, layout.tril())
        out = att(q, k, v, key_padding_mask=kpm)
        dense_mask = layout.bool().repeat_interleave(blk, 1).repeat_interleave(blk, 2)[None]
        sc = (q @ k.transpose(-1, -2)) * D**-0.5 + kpm[:, None, None, :]
        sc = sc.masked_fill(~dense_mask, float("-inf"))
        ref = torch.softmax(sc, -1) @ v
        torch.testing.assert_close(out, ref, atol=1e-5, rtol=1e-4)
    out.sum().backward()
    assert q.grad is not None and torch.isfinite(q.grad).all()


def test_block_sparse_matmul_softmax():
    from deepspeed_b200.ops.sparse_attention import FixedSparsityConfig, MatMul, Softmax
    torch.manual_seed(0)
    B, H, S, D, blk = 2, 2, 32, 8, 8
    layout = FixedSparsityConfig(H, blk, num_local_blocks=2).make_layout(S)
    q, k, v = (torch.randn(B, H, S, D) for _ in range(3))
    sdd = MatMul(layout, blk, "sdd", trans_b=True)
    dsd = MatMul(layout, blk, "dsd")
    sm = Softmax(layout, blk)
    w = sdd(q, k)
    p = sm(w, scale=D**-0.5)
    out = dsd(p, v)
    mask = layout.bool().repeat_interleave(blk, 1).repeat_interleave(blk, 2)[None]
    sc = ((q @ k.transpose(-1, -2)) * D**-0.5).masked_fill(~mask, float("-inf"))
    ref = torch.softmax(sc, -1) @ v
    torch.testing.assert_close(out, ref, atol=1e-5, rtol=1e-4)


def test_evoformer_attention_with_bias_grads():
    from deepspeed_b200.ops.deepspeed4science import DS4Sci_EvoformerAttention
    torch.manual_seed(0)
    Bt, N, L, H, D = 1, 3, 10, 2, 4
    q, k, v = (torch.randn(Bt, N, L, H, D, dtype=torch.float64, requires_grad=True) for _ in range(3))
    mask = torch.randn(Bt, N, 1, 1, L, dtype=torch.float64, requires_grad=True)
    pair = torch.randn(Bt, 1, H, L, L, dtype=torch.float64, requires_grad=True)

    def ref(q, k, v, b1, b2):
        qt, kt, vt = (t.transpose(-2, -3) for t in (q, k, v))
        sc = qt @ kt.transpose(-1, -2) * D**-0.5 + b1 + b2
        return (torch.softmax(sc, -1) @ vt).transpose(-2, -3)

    out = DS4Sci_EvoformerAttention(q, k, v, [mask, pair])
    want = ref(q, k, v, mask, pair)
    torch.testing.assert_close(out, want)
    g = torch.randn_like(out)
    got = torch.autograd.grad(out, (q, k, v, mask, pair), g)
    exp = torch.autograd.grad(want, (q, k, v, mask, pair), g)
    for a, b in zip(got, exp):
        torch.testing.assert_close(a, b, atol=1e-9, rtol=1e-7)


def test_random_ltd_ops():
    from deepspeed_b200.ops.random_ltd import GatherTokens, ScatterTokens, bert_sample_tokens, gpt_sample_tokens
    torch.manual_seed(0)
    idx, _ = gpt_sample_tokens(5, 12, 2, layers=3)
    assert idx.shape == (3, 2, 5) and (idx[..., 1:] > idx[..., :-1]).all()
    x = torch.randn(2, 12, 6, requires_grad=True)
    _, part = GatherTokens.apply(x, idx[0], True)
    assert torch.equal(part, torch.gather(x, 1, idx[0].long()[..., None].expand(2, 5, 6)))
    y = ScatterTokens.apply(x, part * 2, idx[0], True)
    ref = x.detach().clone()
    ref.scatter_(1, idx[0].long()[..., None].expand(2, 5, 6), part.detach() * 2)
    assert torch.equal(y, ref)
    y.sum().backward()
    exp = torch.ones(2, 12, 6)
    exp.scatter_(1, idx[0].long()[..., None].expand(2, 5, 6), torch.full((2, 5, 6), 2.0))
    torch.testing.assert_close(x.grad, exp)
    m = torch.randn(2, 1, 12, 12)
    _, masks = bert_sample_tokens(5, 12, 2, layers=1, attn_mask=m)
    assert masks[0].shape == (2, 1, 5, 5)


def test_misc_host_definitions():
    from deepspeed_b200.ops.kernels import misc_ops as K
    torch.manual_seed(0)
    s = torch.randn(2, 3, 4, 6)
    p = K.attn_softmax(s, scale=0.5, causal=True)
    i = torch.arange(4)[:, None] + 2
    ref = torch.softmax((s * 0.5).masked_fill(torch.arange(6)[None, :] > i, float("-inf")), -1)
    torch.testing.assert_close(p, ref)
    x = torch.randn(2, 5, 3, 2, 4)
    b = torch.randn(3 * 2 * 4)
    t = K.bias_transform_0213(x, b, 2, 5, 3, 2, 4)
    assert t.shape == (3, 2, 2, 5, 4)
    torch.testing.assert_close(t[1, 0, 1, 3], x[0, 3, 1, 1] + b.view(3, 2, 4)[1, 1])
    y = K.dropout(torch.ones(1000, 8), 0.25, training=True)
    assert abs((y == 0).float().mean().item() - 0.25) < 0.05
    from deepspeed_b200.ops.flatten import flatten, unflatten
    ts = [torch.randn(3, 2), torch.randn(5)]
    f = flatten(ts)
    for a, b2 in zip(unflatten(f, ts), ts):
        assert torch.equal(a, b2)


def _shm_worker():
    import torch.distributed as td
    from deepspeed_b200.comm.shm import ShmComm
    r, w = td.get_rank(), td.get_world_size()
    import os
    c = ShmComm(r, w, name=f"/dsb200_test_{os.getppid()}", max_bytes=1 << 16)
    t = torch.full((50_000, ), float(r + 1))
    c.all_reduce(t)
    assert torch.all(t == sum(range(1, w + 1)))
    out = torch.empty(w * 4)
    c.all_gather(out, torch.full((4, ), float(r)))
    assert torch.equal(out, torch.arange(w).float().repeat_interleave(4))
    c.barrier()
    c.close()


def test_shm_collectives():
    run_distributed(_shm_worker, 2)


def test_native_block_allocator_and_atoms():
    from deepspeed_b200.inference.v2.ragged.host import NativeBlockAllocator, build_atoms
    a = NativeBlockAllocator(8)
    x = a.allocate(3)
    assert a.free_blocks == 5
    a.free(x)
    assert a.free_blocks == 8
    with pytest.raises(ValueError):
        a.allocate(9)
    atoms = build_atoms(torch.tensor([5, 1]), torch.tensor([0, 40]), torch.tensor([0, 5]), torch.tensor([0, 8]), 4, 16)
    assert atoms.tolist() == [[0, 0, 4, 1, 4, 0], [0, 4, 1, 1, 5, 0], [1, 5, 1, 3, 41, 8]]
