"""Kernel objects of the ragged engine (inference/v2/kernels/*) vs plain torch, host tier."""
import math

import pytest
import torch

from deepspeed_b200.inference.v2.kernels.core_ops import (BlasLibLinear, CUDABiasActivation, CUDAFPLN, CUDAFPPostLN, CUDAFPPreLN,
                                                          CUDAGatedActivation, CUDARMSNorm, CUDARMSPreNorm)
from deepspeed_b200.inference.v2.kernels.cutlass_ops import MixedGEMM, MoEGEMM
from deepspeed_b200.inference.v2.kernels.ragged_ops import (AtomBuilder, BlockedFlashAttn, BlockedRotaryEmbeddings,
                                                            LinearBlockedKVCopy, MoEGather, MoEScatter, RaggedEmbeddingKernel,
                                                            RaggedLogitsGather, RaggedTopKGating)
from deepspeed_b200.utils.types import ActivationFuncType


def test_core_ops():
    torch.manual_seed(0)
    x, y = torch.randn(6, 32), torch.randn(6, 32)
    g, b = torch.randn(32), torch.randn(32)
    out = torch.empty_like(x)
    ln = torch.nn.functional.layer_norm
    assert torch.allclose(CUDAFPLN(32, torch.float32, 1e-5)(out, x, g, b), ln(x, (32, ), g, b, 1e-5), atol=1e-5)
    assert torch.allclose(CUDAFPPostLN(32, torch.float32)(out, x, y, g, b), ln(x + y, (32, ), g, b, 1e-5), atol=1e-5)
    zr, zh = torch.empty_like(x), torch.empty_like(x)
    CUDAFPPreLN(32, torch.float32)(zr, zh, x, y, g, b)
    assert torch.allclose(zr, x + y, atol=1e-6) and torch.allclose(zh, ln(x + y, (32, ), g, b, 1e-5), atol=1e-5)
    rms = lambda t: t * torch.rsqrt(t.pow(2).mean(-1, keepdim=True) + 1e-5) * g
    assert torch.allclose(CUDARMSNorm(32, torch.float32)(out, x, g), rms(x), atol=1e-5)
    CUDARMSPreNorm(32, torch.float32)(zr, zh, x, y, g)
    assert torch.allclose(zh, rms(x + y), atol=1e-5)
    a = x.clone()
    CUDABiasActivation(32, torch.float32, ActivationFuncType.ReLU)(a, b)
    assert torch.allclose(a, torch.relu(x + b))
    go = torch.empty(6, 16)
    CUDAGatedActivation(32, torch.float32, ActivationFuncType.GATED_SILU)(go, x)
    assert torch.allclose(go, torch.nn.functional.silu(x[:, :16]) * x[:, 16:], atol=1e-6)
    w = torch.randn(8, 32)
    lo = torch.empty(6, 8)
    assert torch.allclose(BlasLibLinear(torch.float32)(lo, x, w), x @ w.t(), atol=1e-5)
    with pytest.raises(ValueError):
        CUDAFPLN(32, torch.int8)
    with pytest.raises(ValueError):
        CUDABiasActivation(30, torch.float16, ActivationFuncType.GELU)


def test_quantised_and_grouped_gemms():
    torch.manual_seed(0)
    mg = MixedGEMM(torch.bfloat16, num_bits=8)
    w = torch.randn(16, 128) * 0.1
    qw = mg.quantize(w.bfloat16(), group_size=64)
    x = torch.randn(3, 128).bfloat16()
    out = torch.empty(3, 16, dtype=torch.bfloat16)
    mg(out, x, qw)
    assert (out.float() - x.float() @ w.t()).abs().max() < 0.05
    xs = torch.randn(7, 8)
    ew = torch.randn(3, 4, 8)
    o = torch.empty(7, 4)
    MoEGEMM(torch.float32)(o, xs, ew, torch.tensor([2, 2, 7]))
    assert torch.allclose(o[:2], xs[:2] @ ew[0].t(), atol=1e-5) and torch.allclose(o[2:], xs[2:] @ ew[2].t(), atol=1e-5)


def test_ragged_attention_pipeline():
    """embed -> rotary + KV append -> blocked attention -> last-token gather on two sequences of a ragged batch."""
    torch.manual_seed(0)
    hq, hkv, d, bs = 4, 2, 16, 4
    lens = [5, 3]
    T = sum(lens)
    seq_of = torch.tensor([0] * 5 + [1] * 3, dtype=torch.int32)
    pos_of = torch.tensor(list(range(5)) + list(range(3)), dtype=torch.int32)
    block_table = torch.tensor([[0, 1], [2, 3]], dtype=torch.int32)
    cache = torch.zeros(4, bs, 2, hkv, d)
    wte = torch.randn(50, 32)
    ids = torch.randint(0, 50, (T, ), dtype=torch.int32)
    emb = torch.empty(T, 32)
    RaggedEmbeddingKernel(torch.float32, torch.int32, 32)(emb, ids, wte)
    assert torch.equal(emb, wte[ids.long()])
    qkv = torch.randn(T, (hq + 2 * hkv) * d)
    ref_qkv = qkv.clone()
    LinearBlockedKVCopy(d, hq, hkv, torch.float32)(cache, qkv, seq_of, pos_of, block_table, bs)
    out = torch.empty(T, hq * d)
    BlockedFlashAttn(d, torch.float32)(out, qkv, cache, seq_of, pos_of, block_table, hq, hkv, bs)
    v = ref_qkv.view(T, hq + 2 * hkv, d)
    start = 0
    for n in lens:
        q = v[start:start + n, :hq].transpose(0, 1)
        k = v[start:start + n, hq:hq + hkv].transpose(0, 1).repeat_interleave(2, 0)
        vv = v[start:start + n, hq + hkv:].transpose(0, 1).repeat_interleave(2, 0)
        ref = torch.nn.functional.scaled_dot_product_attention(q, k, vv, is_causal=True).transpose(0, 1).reshape(n, hq * d)
        assert torch.allclose(out[start:start + n], ref, atol=1e-5)
        start += n
    # rotary variant only changes q/k consistently: attention of a single token (pos 0) is unaffected
    rot = BlockedRotaryEmbeddings(d, hq, hkv, torch.float32, d, 10000.0, max_positions=16)
    q2 = ref_qkv.clone()
    rot(torch.zeros_like(cache), q2, seq_of, pos_of, block_table, bs)
    assert torch.allclose(q2[0], ref_qkv[0], atol=1e-6) and not torch.allclose(q2[1], ref_qkv[1])
    last = torch.empty(2, hq * d)
    RaggedLogitsGather(hq * d, torch.float32)(last, out, torch.tensor([4, 7], dtype=torch.int32))
    assert torch.equal(last, out[[4, 7]])
    atoms, n = AtomBuilder()(torch.zeros(8, 8, dtype=torch.int32), [[0, 5, 0], [5, 3, 0]], q_block_size=4, kv_block_size=bs)
    assert n == 3 and atoms[1].tolist()[:6] == [0, 4, 1, 2, 5, 4]


def test_moe_gating_scatter_gather_roundtrip():
    torch.manual_seed(0)
    T, E, k, H = 6, 4, 2, 8
    logits = torch.randn(T, E)
    counts, scores = torch.zeros(E, dtype=torch.int32), torch.zeros(T, k)
    assign, offs = torch.zeros(T, k, dtype=torch.int32), torch.zeros(T, k, dtype=torch.int32)
    RaggedTopKGating(torch.float32)(counts, scores, assign, offs, logits)
    probs = logits.softmax(-1)
    tv, ti = probs.topk(k, dim=-1)
    assert torch.equal(assign.long().sort(-1)[0], ti.sort(-1)[0]) and int(counts.sum()) == T * k
    x = torch.randn(T, H)
    moe_in, cum, slots = torch.zeros(T * k, H), torch.zeros(E, dtype=torch.int32), torch.zeros(T, k, dtype=torch.int32)
    MoEScatter(torch.float32, H)(moe_in, cum, slots, x, counts, assign, offs)
    assert int(cum[-1]) == T * k
    for t in range(T):
        for j in range(k):
            assert torch.equal(moe_in[int(slots[t, j])], x[t])
    out = torch.empty(T, H)
    MoEGather(torch.float32, H, normalize_scores=False)(out, moe_in * 2.0, scores, slots)
    assert torch.allclose(out, x * 2.0 * scores.sum(-1, keepdim=True), atol=1e-5)
