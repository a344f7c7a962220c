"""Small standalone launches of the framework's hot kernels, for `ncu --set full` captures.
Usage: python scripts/ncu_targets.py adam|rmsnorm|swiglu|xent|rope"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from deepspeed_b200.ops.kernels import flat_ops, transformer_ops as T

which = sys.argv[1] if len(sys.argv) > 1 else "adam"
d = "cuda"
torch.manual_seed(0)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=d)
if which == "attn":
    from deepspeed_b200.ops.kernels import attention_sm100 as A
    B, S, hq, hkv = 2, 4096, 32, 8
    qkv = torch.randn(B * S, (hq + 2 * hkv) * 128, device=d, dtype=torch.bfloat16)
    q, k, v = A.split_packed(qkv, hq, hkv)
    for _ in range(2):
        o, lse = A.fwd(q, k, v, B, S, hq, hkv, causal=True)
        A.bwd(torch.randn_like(o), q, k, v, o, lse, B, S, hq, hkv, causal=True)
    torch.cuda.synchronize()
elif which == "evo":
    from deepspeed_b200.ops.kernels import attn_bias as AB
    NB, H, L, D = 128, 8, 256, 32
    q, k, v = (torch.randn(NB, L, H, D, device=d, dtype=torch.bfloat16).permute(0, 2, 1, 3) for _ in range(3))
    b1 = torch.zeros(NB, L, device=d, dtype=torch.bfloat16)
    b2 = torch.randn(1, H, L, L, device=d, dtype=torch.bfloat16)
    for _ in range(2):
        o, lse = AB.forward(q, k, v, b1, b2)
        AB.backward(torch.randn_like(o), q, k, v, o, lse, b1, b2, need_db1=True, need_db2=True)
    torch.cuda.synchronize()
elif which == "mlp":
    from deepspeed_b200.ops import gemm as G
    x = torch.randn(8192, 4096, device=d, dtype=torch.bfloat16)
    w_gu = torch.randn(2 * 14336, 4096, device=d, dtype=torch.bfloat16) * 0.02
    w_dn = torch.randn(4096, 14336, device=d, dtype=torch.bfloat16) * 0.02
    for _ in range(2):
        flush.zero_()
        act, gu = G.gate_up_swiglu(x, w_gu, save_gate_up=True)
        dy = torch.randn(8192, 4096, device=d, dtype=torch.bfloat16)
        G.down_dx_dswiglu(dy, w_dn, gu)
    torch.cuda.synchronize()
elif which == "wqtc":
    from deepspeed_b200.inference.quantization import layers as QL
    w = torch.randn(14336, 4096, device=d, dtype=torch.bfloat16) * 0.02
    x = torch.randn(256, 4096, device=d, dtype=torch.bfloat16)
    for mode in ("int8", "fp6"):
        qw = QL.quantize_weight(w, mode=mode, group_size=128)
        for _ in range(2):
            flush.zero_()
            QL.wq_tc_linear(x, qw)
    torch.cuda.synchronize()
elif which == "adam":
    n = 256 * 1024 * 1024
    p = torch.randn(n, device=d)
    g = torch.randn(n, device=d, dtype=torch.bfloat16)
    m, v = torch.zeros(n, device=d), torch.zeros(n, device=d)
    out = torch.empty(n, device=d, dtype=torch.bfloat16)
    for step in range(1, 6):
        flush.zero_()
        flat_ops.adam_flat(p, g, m, v, out, lr=1e-4, beta1=0.9, beta2=0.95, eps=1e-8, weight_decay=0.1, step=step)
elif which == "rmsnorm":
    x = torch.randn(8192, 4096, device=d, dtype=torch.bfloat16, requires_grad=True)
    r = torch.randn(8192, 4096, device=d, dtype=torch.bfloat16, requires_grad=True)
    w = torch.ones(4096, device=d, dtype=torch.bfloat16, requires_grad=True)
    for _ in range(5):
        flush.zero_()
        y, s = T.rms_norm(x, w, 1e-5, residual=r)
        (y.sum() + s.sum()).backward()
elif which == "swiglu":
    gu = torch.randn(8192, 2 * 14336, device=d, dtype=torch.bfloat16, requires_grad=True)
    for _ in range(5):
        flush.zero_()
        T.gated_act(gu, "silu").sum().backward()
elif which == "xent":
    lg = torch.randn(2048, 128256, device=d, dtype=torch.bfloat16)
    lab = torch.randint(0, 128256, (2048, ), device=d)
    for _ in range(5):
        flush.zero_()
        T.softmax_xent_fwd_bwd(lg.clone(), lab, 1.0)
elif which == "rope":
    table = T.RotaryTable(128, 8192, 500000.0, d)
    qkv = torch.randn(8192, 6144, device=d, dtype=torch.bfloat16)
    for _ in range(5):
        flush.zero_()
        T.rope_qk_inplace(qkv, 32, 8, 128, table, None, 4096)
elif which == "gemm":
    from deepspeed_b200.ops.kernels import gemm_sm100
    a = torch.randn(8192, 4096, device=d, dtype=torch.bfloat16)
    b = torch.randn(14336 * 2, 4096, device=d, dtype=torch.bfloat16)
    for _ in range(3):
        flush.zero_()
        gemm_sm100.matmul_nt(a, b)
elif which == "gemm2cta":
    from deepspeed_b200.ops.kernels import gemm_sm100
    a = torch.randn(8192, 4096, device=d, dtype=torch.bfloat16)
    b = torch.randn(6144, 4096, device=d, dtype=torch.bfloat16)
    for _ in range(3):
        flush.zero_()
        gemm_sm100.matmul_nt_2cta(a, b)
elif which == "paged":
    from deepspeed_b200.ops.kernels import ragged_ops as R
    hq, hkv, dd, bs, seqs, ctx = 32, 8, 128, 128, 64, 2048
    qkv = torch.randn(seqs, (hq + 2 * hkv) * dd, device=d, dtype=torch.bfloat16)
    nb = ctx // bs
    cache = torch.randn(seqs * nb, bs, 2, hkv, dd, device=d, dtype=torch.bfloat16)
    bt = torch.arange(seqs * nb, device=d, dtype=torch.int32).view(seqs, nb)
    seq_of = torch.arange(seqs, device=d, dtype=torch.int32)
    pos_of = torch.full((seqs, ), ctx - 1, device=d, dtype=torch.int32)
    for _ in range(3):
        flush.zero_()
        R.paged_attention(qkv, cache, seq_of, pos_of, bt, hq, hkv, dd, bs)
elif which == "grouped":
    from deepspeed_b200.ops.kernels import gemm_sm100
    E, N, K = 8, 28672, 4096  # Mixtral gate+up projection
    w = (torch.randn(E, N, K, device=d) * 0.02).bfloat16()
    for rows_per in (2, 512):  # decode-like (weight streaming) and prefill-like (tensor-core bound)
        x = (torch.randn(E * rows_per, K, device=d) * 0.5).bfloat16()
        off = (torch.arange(E + 1, device=d, dtype=torch.int32) * rows_per)
        for _ in range(2):
            flush.zero_()
            gemm_sm100.grouped_matmul_nt(x, w, off)
elif which == "wq":
    from deepspeed_b200.inference.quantization.layers import maybe_quantized_linear, quantize_weight
    N, K, M = 28672, 4096, 8
    w = (torch.randn(N, K, device=d) * 0.05).bfloat16()
    x = torch.randn(M, K, device=d).bfloat16()
    for mode in ("int8", "int4"):
        qw = quantize_weight(w, mode, 128)
        for _ in range(2):
            flush.zero_()
            maybe_quantized_linear(x, qw)
torch.cuda.synchronize()
print("done", which)
