"""Evoformer (AlphaFold-style) attention with up to two additive biases (reference
``ops/deepspeed4science/evoformer_attn.py`` over the CUTLASS kernels in ``csrc/deepspeed4science`` N12).

Inputs ``Q/K/V [*, L, H, D]`` with arbitrary leading dims (MSA row / column, pair), ``biases``: a mask bias
broadcast as ``[*, 1, 1, L]`` and a pair bias ``[1.., H, L, L]``.  The fused flash kernel (SDPA with an additive
mask) does the work; the bias gradients — the part the reference needs a custom backward for — come from a
custom autograd function that recomputes the probabilities chunk-wise so the [*, H, L, L] score tensor is never
materialised for more than ``chunk`` leading rows at a time.
"""
import torch
import torch.nn.functional as F


def _flat(x):
    return x.reshape(-1, *x.shape[-3:])


class EvoformerFusedAttention(torch.autograd.Function):

    @staticmethod
    def forward(ctx, q, k, v, bias1=None, bias2=None, chunk=64):
        # [*, L, H, D] -> [*, H, L, D]
        qt, kt, vt = (t.transpose(-2, -3) for t in (q, k, v))
        lead = qt.shape[:-3]
        bias = None
        for b in (bias1, bias2):
            if b is not None:
                bias = b if bias is None else bias + b
        if bias is not None:
            bias = bias.expand(*lead, *bias.shape[-3:]) if bias.dim() == qt.dim() else bias
        o = F.scaled_dot_product_attention(qt, kt, vt, attn_mask=bias.to(qt.dtype) if bias is not None else None)
        ctx.save_for_backward(q, k, v, bias1, bias2, o)
        ctx.chunk = chunk
        return o.transpose(-2, -3).contiguous()

    @staticmethod
    def backward(ctx, do):
        q, k, v, bias1, bias2, o = ctx.saved_tensors
        dq, dk, dv, db1, db2 = attention_bwd(do, q, k, v, o.transpose(-2, -3), None, bias1, bias2,
                                             bias1 is not None and ctx.needs_input_grad[3],
                                             bias2 is not None and ctx.needs_input_grad[4], chunk=ctx.chunk)
        return dq, dk, dv, db1, db2, None


def attention_bwd(dO, Q, K, V, O, lse, bias1, bias2, bias1_grad, bias2_grad, chunk=64):
    """Gradients of evoformer attention for ``[*, L, H, D]`` operands (reference ``evoformer_attn.py:33``).  The softmax is
    recomputed chunk by chunk over the flattened leading dims (``lse`` is accepted for signature parity and not needed);
    returns ``(dQ, dK, dV, dB1, dB2)`` with ``None`` for bias gradients that were not requested."""
    qt, kt, vt, ot, dot = (t.transpose(-2, -3) for t in (Q, K, V, O, dO))
    shape = qt.shape
    Qf, Kf, Vf, Of, DOf = (_flat(t) for t in (qt, kt, vt, ot, dot))
    N, H, L, D = Qf.shape
    scale = D**-0.5
    ct = torch.float64 if Qf.dtype == torch.float64 else torch.float32
    b1 = _flat(bias1.expand(*shape[:-3], *bias1.shape[-3:])) if bias1 is not None else None
    b2 = _flat(bias2.expand(*shape[:-3], *bias2.shape[-3:])) if bias2 is not None else None
    dQ, dK, dV = torch.empty_like(Qf), torch.empty_like(Kf), torch.empty_like(Vf)
    db1 = torch.zeros(bias1.shape, dtype=ct, device=Q.device) if (bias1 is not None and bias1_grad) else None
    db2 = torch.zeros(bias2.shape, dtype=ct, device=Q.device) if (bias2 is not None and bias2_grad) else None
    for s in range(0, N, chunk):
        e = min(s + chunk, N)
        sc = torch.matmul(Qf[s:e].to(ct), Kf[s:e].to(ct).transpose(-1, -2)) * scale
        if b1 is not None:
            sc = sc + b1[s:e].to(ct)
        if b2 is not None:
            sc = sc + b2[s:e].to(ct)
        P = torch.softmax(sc, -1)
        dOc = DOf[s:e].to(ct)
        dV[s:e] = torch.matmul(P.transpose(-1, -2), dOc).to(Vf.dtype)
        dP = torch.matmul(dOc, Vf[s:e].to(ct).transpose(-1, -2))
        delta = (dOc * Of[s:e].to(ct)).sum(-1, keepdim=True)
        dS = P * (dP - delta)
        dQ[s:e] = (torch.matmul(dS, Kf[s:e].to(ct)) * scale).to(Qf.dtype)
        dK[s:e] = (torch.matmul(dS.transpose(-1, -2), Qf[s:e].to(ct)) * scale).to(Kf.dtype)
        if db1 is not None:
            _accum_bias_grad(db1, dS, bias1, shape, s, e)
        if db2 is not None:
            _accum_bias_grad(db2, dS, bias2, shape, s, e)
    un = lambda t: t.view(shape).transpose(-2, -3)
    return (un(dQ), un(dK), un(dV), db1.to(bias1.dtype) if db1 is not None else None,
            db2.to(bias2.dtype) if db2 is not None else None)


def _accum_bias_grad(db, dS, bias, full_shape, s, e):
    """Reduce dS [chunk, H, L, L] into ``db`` (shape of ``bias`` incl. broadcast dims)."""
    lead = full_shape[:-3]
    n_lead = 1
    for x in lead:
        n_lead *= x
    # position of rows s..e in the flattened leading index -> unravel and scatter-add over broadcast dims
    idx = torch.arange(s, e, device=dS.device)
    blead = bias.shape[:-3]
    # align bias leading dims to full leading dims from the right
    pad = len(lead) - len(blead)
    strides, mult = [], 1
    for d in reversed(range(len(lead))):
        strides.append(mult)
        mult *= lead[d]
    strides = list(reversed(strides))
    tgt = torch.zeros_like(idx)
    bmult = 1
    for d in reversed(range(len(lead))):
        coord = (idx // strides[d]) % lead[d]
        bd = d - pad
        if bd >= 0 and blead[bd] != 1:
            tgt = tgt + coord * bmult
            bmult *= blead[bd]
    red = dS
    for ax, n in zip((1, 2, 3), bias.shape[-3:]):
        if n == 1:
            red = red.sum(ax, keepdim=True)
    db.view(-1, *bias.shape[-3:]).index_add_(0, tgt, red)


def DS4Sci_EvoformerAttention(Q, K, V, biases):
    assert len(biases) <= 2
    biases = list(biases) + [None] * (2 - len(biases))
    b1, b2 = biases
    if b1 is not None:
        assert b1.shape[-3] == 1 and b1.shape[-2] == 1, "bias1 is the mask bias: shape [*, 1, 1, L]"
    if b2 is not None:
        assert b2.shape[-3] == Q.shape[-2], "bias2 is the pair bias: shape [*, H, L, L]"
    return EvoformerFusedAttention.apply(Q, K, V, b1, b2)
