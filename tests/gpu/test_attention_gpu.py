"""GPU tier: the tcgen05 training attention kernel against an fp32 reference."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(q, k, v, causal):
    """fp32 reference: q [B, hq, S, d], k/v [B, hkv, S, d]."""
    B, hq, S, d = q.shape
    rep = hq // k.shape[1]
    k, v = k.repeat_interleave(rep, 1), v.repeat_interleave(rep, 1)
    s = (q.float() @ k.float().transpose(-1, -2)) / math.sqrt(d)
    if causal:
        s = s.masked_fill(torch.triu(torch.ones(S, S, dtype=torch.bool, device=q.device), 1), float("-inf"))
    return torch.softmax(s, -1) @ v.float(), torch.logsumexp(s, -1)


@pytest.mark.parametrize("B,S,hq,hkv,causal", [(1, 128, 1, 1, True), (1, 256, 2, 1, True), (2, 512, 4, 2, True),
                                               (1, 384, 4, 4, False), (1, 1024, 8, 2, True)])
def test_attn_fwd_matches_fp32(B, S, hq, hkv, causal):
    from deepspeed_b200.ops.kernels import attention_sm100 as A
    torch.manual_seed(S + hq)
    d = 128
    qkv = torch.randn(B * S, (hq + 2 * hkv) * d, device="cuda", dtype=torch.bfloat16)
    q2, k2, v2 = A.split_packed(qkv, hq, hkv)
    assert A.supports(qkv, hq, hkv, d, S)
    o, lse = A.fwd(q2, k2, v2, B, S, hq, hkv, causal=causal)
    torch.cuda.synchronize()
    x = qkv.view(B, S, hq + 2 * hkv, d)
    ref_o, ref_lse = _ref(x[:, :, :hq].transpose(1, 2), x[:, :, hq:hq + hkv].transpose(1, 2),
                          x[:, :, hq + hkv:].transpose(1, 2), causal)
    got = o.view(B, S, hq, d).transpose(1, 2).float()
    assert torch.isfinite(got).all()
    assert (got - ref_o).abs().max().item() < 2e-2, (got - ref_o).abs().max().item()
    assert (lse - ref_lse).abs().max().item() < 2e-2
    # repeated launches: barrier phases / TMEM reuse stay correct
    o2, _ = A.fwd(q2, k2, v2, B, S, hq, hkv, causal=causal)
    assert torch.equal(o, o2)
