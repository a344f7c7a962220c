"""Biased / block-sparse flash attention kernels (csrc/cuda/attn_bias.cu) against an fp32 PyTorch reference: forward,
dQ / dK / dV and both bias gradients; Evoformer and sparse-attention front ends."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(q, k, v, b1, b2, b2_div, mask, scale):
    """fp32 reference on [NB, H, L, D] operands; returns out and grads via autograd."""
    s = torch.matmul(q, k.transpose(-1, -2)) * scale
    if b1 is not None:
        s = s + b1[:, None, None, :]
    if b2 is not None:
        s = s + b2.repeat_interleave(b2_div, dim=0)
    if mask is None:
        return torch.matmul(torch.softmax(s, -1), v)
    mask = mask.expand(s.shape)
    dead = ~mask.any(-1, keepdim=True)  # rows without any visible key: zero output, zero gradient
    p = torch.softmax(s.masked_fill(~mask & ~dead, float("-inf")), -1).masked_fill(dead, 0.0)
    return torch.matmul(p, v)


def _mk(NB, H, Lq, Lk, D, dtype, layout_evo=True):
    g = torch.Generator(device="cuda").manual_seed(NB * 1000 + Lq + D)
    if layout_evo:  # [NB, L, H, D] memory, logical [NB, H, L, D]
        mk = lambda L: torch.randn(NB, L, H, D, device="cuda", dtype=dtype, generator=g).permute(0, 2, 1, 3)
    else:
        mk = lambda L: torch.randn(NB, H, L, D, device="cuda", dtype=dtype, generator=g)
    return mk(Lq), mk(Lk), mk(Lk), g


def _check(got, want, tol, what):
    err = (got.float() - want.float()).abs().max().item()
    ref = want.float().abs().max().item() + 1e-6
    assert err <= tol * ref + tol, f"{what}: max err {err:.4g} vs scale {ref:.4g}"


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("D", [16, 32, 64])
@pytest.mark.parametrize("L", [64, 100, 257])
def test_biased_attention_matches_reference(dtype, D, L):
    from deepspeed_b200.ops.kernels import attn_bias as AB
    B, Nrep, H = 2, 3, 4
    NB = B * Nrep
    q, k, v, g = _mk(NB, H, L, L, D, dtype)
    b1 = (torch.randn(NB, L, device="cuda", generator=g) * 2).to(dtype)
    b1[:, -3:] = -1e4 if dtype == torch.bfloat16 else -3e4  # masked keys, like the Evoformer mask bias
    b2 = torch.randn(B, H, L, L, device="cuda", generator=g).to(dtype)
    qs, ks, vs, b1s, b2s = (t.detach().clone().requires_grad_(True) for t in (q, k, v, b1, b2))
    out = AB.biased_attention(qs, ks, vs, b1s, b2s)
    d_o = torch.randn(out.shape, device="cuda", generator=g).to(dtype)
    out.backward(d_o)
    qf, kf, vf, b1f, b2f = (t.detach().float().requires_grad_(True) for t in (q, k, v, b1, b2))
    ref = _ref(qf, kf, vf, b1f, b2f, Nrep, None, 1 / math.sqrt(D))
    ref.backward(d_o.float())
    tol = 2e-2 if dtype == torch.bfloat16 else 4e-3
    _check(out, ref, tol, "out")
    _check(qs.grad, qf.grad, 2 * tol, "dq")
    _check(ks.grad, kf.grad, 2 * tol, "dk")
    _check(vs.grad, vf.grad, 2 * tol, "dv")
    _check(b1s.grad, b1f.grad, 3 * tol, "db1")
    _check(b2s.grad, b2f.grad, 3 * tol, "db2")


@pytest.mark.parametrize("Lq,Lk", [(48, 200), (130, 70)])
def test_cross_lengths_no_bias_and_causal(Lq, Lk):
    from deepspeed_b200.ops.kernels import attn_bias as AB
    q, k, v, g = _mk(3, 2, Lq, Lk, 32, torch.bfloat16, layout_evo=False)
    out, lse = AB.forward(q, k, v)
    ref = _ref(q.float(), k.float(), v.float(), None, None, 1, None, 1 / math.sqrt(32))
    _check(out, ref, 2e-2, "out")
    s = torch.matmul(q.float(), k.float().transpose(-1, -2)) / math.sqrt(32)
    _check(lse, torch.logsumexp(s, -1), 2e-2, "lse")
    # causal (square only)
    q, k, v, g = _mk(2, 2, 150, 150, 64, torch.bfloat16, layout_evo=False)
    qs, ks, vs = (t.detach().clone().requires_grad_(True) for t in (q, k, v))
    out = AB.biased_attention(qs, ks, vs, None, None, None, 0, True, None)
    out.float().square().sum().backward()
    qf, kf, vf = (t.detach().float().requires_grad_(True) for t in (q, k, v))
    mask = torch.ones(150, 150, device="cuda", dtype=torch.bool).tril()
    ref = _ref(qf, kf, vf, None, None, 1, mask, 1 / 8)
    ref.square().sum().backward()
    _check(out, ref, 2e-2, "causal out")
    _check(qs.grad, qf.grad, 5e-2, "causal dq")
    _check(ks.grad, kf.grad, 5e-2, "causal dk")
    _check(vs.grad, vf.grad, 5e-2, "causal dv")


@pytest.mark.parametrize("block", [16, 32, 64, 128])
def test_block_sparse_layout(block):
    from deepspeed_b200.ops.kernels import attn_bias as AB
    B, H, S, D = 2, 4, 512, 64
    q, k, v, g = _mk(B, H, S, S, D, torch.bfloat16, layout_evo=False)
    nb = S // block
    layout = (torch.rand(H, nb, nb, device="cuda", generator=g) < 0.3)
    layout |= torch.eye(nb, device="cuda", dtype=torch.bool)[None]
    layout[0, 1] = False  # a query block row with NO visible key block: zero output, zero gradients
    mask = layout.repeat_interleave(block, 1).repeat_interleave(block, 2)[None]
    qs, ks, vs = (t.detach().clone().requires_grad_(True) for t in (q, k, v))
    out = AB.biased_attention(qs, ks, vs, None, None, layout.to(torch.uint8), block, False, None)
    d_o = torch.randn(out.shape, device="cuda", generator=g).to(torch.bfloat16)
    out.backward(d_o)
    qf, kf, vf = (t.detach().float().requires_grad_(True) for t in (q, k, v))
    ref = _ref(qf, kf, vf, None, None, 1, mask, 1 / 8)
    ref.backward(d_o.float())
    _check(out, ref, 2e-2, "out")
    _check(qs.grad, qf.grad, 4e-2, "dq")
    _check(ks.grad, kf.grad, 4e-2, "dk")
    _check(vs.grad, vf.grad, 4e-2, "dv")
    assert out[:, 0, block:2 * block].abs().max().item() == 0.0


def test_evoformer_front_end_uses_native_kernels():
    from deepspeed_b200.ops import native as N
    from deepspeed_b200.ops.deepspeed4science import DS4Sci_EvoformerAttention
    torch.manual_seed(0)
    B, Nseq, L, H, D = 1, 16, 96, 4, 32
    q, k, v = (torch.randn(B, Nseq, L, H, D, device="cuda", dtype=torch.bfloat16, requires_grad=True) for _ in range(3))
    mask = (torch.rand(B, Nseq, 1, 1, L, device="cuda") < 0.1).to(torch.bfloat16) * -1e4
    mask.requires_grad_(True)
    pair = torch.randn(B, 1, H, L, L, device="cuda", dtype=torch.bfloat16, requires_grad=True)
    n0 = N.launch_count
    out = DS4Sci_EvoformerAttention(q, k, v, [mask, pair])
    d_o = torch.randn_like(out)
    out.backward(d_o)
    assert N.launch_count >= n0 + 2, "the native forward + backward entry points must have run"
    qf, kf, vf, mf, pf = (t.detach().float().requires_grad_(True) for t in (q, k, v, mask, pair))
    s = torch.einsum("bnqhd,bnkhd->bnhqk", qf, kf) / math.sqrt(D) + mf + pf
    ref = torch.einsum("bnhqk,bnkhd->bnqhd", torch.softmax(s, -1), vf)
    ref.backward(d_o.float())
    _check(out, ref, 2e-2, "out")
    for a, b, nm in ((q, qf, "dq"), (k, kf, "dk"), (v, vf, "dv"), (mask, mf, "dmask"), (pair, pf, "dpair")):
        assert a.grad.shape == b.grad.shape
        _check(a.grad, b.grad, 5e-2, nm)


def test_sparse_self_attention_module_matches_gather_path():
    from deepspeed_b200.ops.sparse_attention import FixedSparsityConfig, SparseSelfAttention
    from deepspeed_b200.ops.sparse_attention import sparse_self_attention as SSA
    torch.manual_seed(0)
    B, H, S, D = 2, 4, 256, 64
    cfg = FixedSparsityConfig(num_heads=H, block=16, num_local_blocks=4, num_global_blocks=1)
    attn = SparseSelfAttention(cfg, max_seq_length=S).cuda()
    q, k, v = (torch.randn(B, H, S, D, device="cuda", dtype=torch.bfloat16) for _ in range(3))
    kpm = torch.zeros(B, S, device="cuda")
    kpm[:, -20:] = -10000.0
    got = attn(q, k, v, key_padding_mask=kpm)
    layout = attn.get_layout(S).cuda()
    orig = SSA._native_ok
    SSA._native_ok = lambda *a: False
    try:
        want = SSA.block_sparse_attention(q, k, v, layout, 16, D**-0.5, kpm, None, "add", "mul")
    finally:
        SSA._native_ok = orig
    _check(got, want, 3e-2, "sparse module")
