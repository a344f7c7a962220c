"""Evoformer (AlphaFold-style) attention with up to two additive biases (reference
``ops/deepspeed4science/evoformer_attn.py`` over the CUTLASS kernels in ``csrc/deepspeed4science`` N12).

Inputs ``Q/K/V [*, L, H, D]`` with arbitrary leading dims (MSA row / column, pair), ``biases``: a mask bias
broadcast as ``[*, 1, 1, L]`` and a pair bias ``[1.., H, L, L]``.

On a GPU with bf16 / fp16 operands and head dim 16 / 32 / 64 (every Evoformer configuration the reference kernel accepts:
``evoformer_attn.py:36`` asserts ``D <= 64``) the work is done by the hand-written biased flash-attention kernels of
``csrc/cuda/attn_bias.cu``: forward with both biases folded into the online softmax, backward producing dQ / dK / dV and
both bias gradients (fp32 accumulation) without ever materialising the ``[*, H, L, L]`` scores.  Other dtypes / head dims
and the CPU use the chunked PyTorch formulation below (SDPA forward, chunk-wise recomputed backward).
"""
import torch
import torch.nn.functional as F


def _flat(x):
    return x.reshape(-1, *x.shape[-3:])


def _native_plan(q, k, v, bias1, bias2):
    """Arguments for the native kernels, or None when the configuration needs the PyTorch path."""
    from deepspeed_b200.ops.kernels import attn_bias as AB
    if not (q.is_cuda and q.dtype in (torch.bfloat16, torch.float16) and q.shape[-1] in AB.HEAD_DIMS and q.dim() >= 4
            and k.shape == q.shape and v.shape == q.shape):
        return None
    lead = q.shape[:-3]
    L, H, D = q.shape[-3:]
    nb = 1
    for x in lead:
        nb *= x
    if nb > AB.MAX_GRID:
        return None
    to4 = lambda t: t.reshape(nb, L, H, D).permute(0, 2, 1, 3)  # logical [NB, H, L, D], no copy for contiguous inputs
    b1 = b2 = None
    if bias1 is not None:
        if bias1.shape[-3:-1] != (1, 1) or bias1.shape[-1] != L:
            return None
        b1 = bias1.expand(*lead, 1, 1, L).reshape(nb, L)
    if bias2 is not None:
        # pair bias [lead' , H, L, L] where lead' broadcasts over the trailing leading dims (N of [B, N]): the kernel shares
        # one bias2 batch entry between `nb / B2` consecutive attention batches
        bl = bias2.shape[:-3]
        if len(bl) != len(lead) or tuple(bias2.shape[-3:]) != (H, L, L):
            return None
        n_b2, seen_one = 1, False
        for have, want in zip(bl, lead):
            if have == want and not seen_one:
                n_b2 *= have
            elif have == 1:
                seen_one = True
            else:
                return None
        b2 = bias2.reshape(n_b2, H, L, L)
    return to4(q), to4(k), to4(v), b1, b2


class EvoformerFusedAttention(torch.autograd.Function):

    @staticmethod
    def forward(ctx, q, k, v, bias1=None, bias2=None, chunk=64):
        plan = _native_plan(q, k, v, bias1, bias2)
        ctx.native = plan is not None
        if plan is not None:
            from deepspeed_b200.ops.kernels import attn_bias as AB
            q4, k4, v4, b1, b2 = plan
            o4, lse = AB.forward(q4, k4, v4, b1, b2)
            ctx.save_for_backward(q, k, v, bias1, bias2, o4, lse)
            return o4.permute(0, 2, 1, 3).reshape(q.shape)
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
        if ctx.native:
            from deepspeed_b200.ops.kernels import attn_bias as AB
            q, k, v, bias1, bias2, o4, lse = ctx.saved_tensors
            q4, k4, v4, b1, b2 = _native_plan(q, k, v, bias1, bias2)
            L, H, D = q.shape[-3:]
            do4 = do.reshape(-1, L, H, D).permute(0, 2, 1, 3)
            dq, dk, dv, db1, db2 = AB.backward(do4, q4, k4, v4, o4, lse, b1, b2, need_db1=b1 is not None and ctx.needs_input_grad[3],
                                               need_db2=b2 is not None and ctx.needs_input_grad[4])
            un = lambda t: t.permute(0, 2, 1, 3).reshape(q.shape)
            if db1 is not None:  # sum over the leading dims bias1 broadcasts over, back to its own shape
                db1 = db1.view(*q.shape[:-3], 1, 1, L).sum_to_size(bias1.shape).to(bias1.dtype)
            if db2 is not None:
                db2 = db2.view(bias2.shape).to(bias2.dtype)
            return un(dq), un(dk), un(dv), db1, db2, None
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
