"""Small single-GPU launches of the hand-synchronised kernels for `compute-sanitizer` (memcheck / racecheck / synccheck):

    compute-sanitizer --tool memcheck python scripts/sanitize_targets.py gemm|attn|wqtc|symm

`symm` drives the NVLink collective kernels in a world-size-1 loop-back (the peer table holds this GPU's own buffers, the
flag pads are plain device memory): same code path, same st.release.sys / ld.acquire.sys flag barriers and grid-exit
protocol as the multi-GPU run, minus the remote mappings.  Every target also checks its result against torch."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from deepspeed_b200.ops import native as N

which = sys.argv[1] if len(sys.argv) > 1 else "gemm"
d = "cuda"
torch.manual_seed(0)


def close(a, b, tol=3e-2):
    a, b = a.float(), b.float()
    assert torch.isfinite(a).all(), "non-finite output"
    assert (a - b).abs().max().item() <= tol * b.abs().max().item() + tol, (a - b).abs().max().item()


if which == "gemm":
    from deepspeed_b200.ops.kernels import gemm_sm100 as K
    M, Nn, Kd, I = 520, 384, 192, 256
    a = torch.randn(M, Kd, device=d).bfloat16()
    b = torch.randn(Nn, Kd, device=d).bfloat16()
    close(K.matmul_2cta(a, b, False, False, group_m=3), a.float() @ b.float().t())
    bk = torch.randn(Kd, Nn, device=d).bfloat16()
    close(K.matmul_2cta(a, bk, False, True), a.float() @ bk.float())
    ak = torch.randn(Kd, M, device=d).bfloat16()
    c0 = torch.randn(M, Nn, device=d).bfloat16()
    c = c0.clone()
    K.matmul_2cta(ak, bk, True, True, out=c, epi=K.EPI_ACCUM)
    close(c, c0.float() + ak.float().t() @ bk.float())
    w = (torch.randn(2 * I, Kd, device=d) * 0.1).bfloat16()
    act = torch.empty(M, I, device=d, dtype=torch.bfloat16)
    gu = torch.empty(M, 2 * I, device=d, dtype=torch.bfloat16)
    K.matmul_2cta(a, w, False, False, out=act, epi=K.EPI_SWIGLU, out2=gu, inter=I)
    r = a.float() @ w.float().t()
    close(gu, r)
    close(act, torch.nn.functional.silu(r[:, :I]) * r[:, I:])
    dy = torch.randn(M, Kd, device=d).bfloat16()
    wd = (torch.randn(Kd, I, device=d) * 0.1).bfloat16()
    dgu = torch.empty(M, 2 * I, device=d, dtype=torch.bfloat16)
    K.matmul_2cta(dy, wd, False, True, out=dgu[:, :I], epi=K.EPI_DSWIGLU, aux=gu, out2=dgu, inter=I)
    g = gu[:, :I].float().requires_grad_(True)
    u = gu[:, I:].float().requires_grad_(True)
    (torch.nn.functional.silu(g) * u).backward(dy.float() @ wd.float())
    close(dgu, torch.cat([g.grad, u.grad], 1))
    x = torch.randn(300, Kd, device=d).bfloat16()
    wg = (torch.randn(3, Nn, Kd, device=d) * 0.1).bfloat16()
    off = torch.tensor([0, 130, 130, 300], dtype=torch.int32, device=d)
    o = K.grouped_matmul_nt(x, wg, off)
    close(o[:130], x[:130].float() @ wg[0].float().t())
    close(o[130:], x[130:].float() @ wg[2].float().t())
elif which == "attn":
    from deepspeed_b200.ops.kernels import attention_sm100 as A
    B, S, hq, hkv = 1, 256, 2, 1
    qkv = (torch.randn(B * S, (hq + 2 * hkv) * 128, device=d) * 0.7).bfloat16()
    q, k, v = A.split_packed(qkv, hq, hkv)
    o, lse = A.fwd(q, k, v, B, S, hq, hkv, causal=True)
    dq, dk, dv = A.bwd(torch.randn_like(o), q, k, v, o, lse, B, S, hq, hkv, causal=True)
    x = qkv.view(B, S, hq + 2 * hkv, 128)
    ref = torch.nn.functional.scaled_dot_product_attention(x[:, :, :hq].transpose(1, 2).float(),
                                                           x[:, :, hq:hq + hkv].transpose(1, 2).float().repeat_interleave(2, 1),
                                                           x[:, :, hq + hkv:].transpose(1, 2).float().repeat_interleave(2, 1),
                                                           is_causal=True)
    close(o.view(B, S, hq, 128).transpose(1, 2), ref)
    o2, _ = A.fwd(q[:200], k[:200], v[:200], 1, 200, hq, hkv, causal=True, need_lse=False)  # ragged length
    assert torch.isfinite(o2.float()).all() and torch.isfinite(dq.float()).all() and torch.isfinite(dk.float()).all()
elif which == "wqtc":
    from deepspeed_b200.inference.quantization.layers import quantize_weight, wq_tc_linear
    for mode in ("fp6", "int4", "fp8", "int8"):
        w = (torch.randn(256, 512, device=d) * 0.05).bfloat16()
        x = (torch.randn(72, 512, device=d) * 0.5).bfloat16()
        qw = quantize_weight(w, mode, 128)
        close(wq_tc_linear(x, qw), x.float() @ qw.dequantize().float().t())
elif which == "symm":
    lib = N.cuda()
    pads = torch.zeros(lib.dsb_symm_pad_bytes() // 4 + 64, dtype=torch.int32, device=d)
    P = lambda t: (ctypes.c_void_p * 1)(ctypes.c_void_p(t.data_ptr()))
    st = N.stream()
    ep = 1
    rc = lib.dsb_symm_barrier(P(pads), 0, 1, 0, ctypes.c_uint32(ep), st); assert rc == 0, rc
    S = 1 << 16
    g = torch.randn(S, device=d).bfloat16()
    for dt_, acc in ((torch.float32, 0), (torch.float32, 1), (torch.bfloat16, 0)):
        dst = torch.ones(S, device=d, dtype=dt_)
        ep += 2
        rc = lib.dsb_symm_reduce_scatter_acc(P(g), ctypes.c_void_p(0), N.ptr(dst), ctypes.c_int64(S), N.dt(g), N.dt(dst),
                                             N.c_f(0.5), acc, P(pads), 0, 1, 2, ctypes.c_uint32(ep), ctypes.c_void_p(0), 16, st)
        assert rc == 0, rc
        close(dst, 0.5 * g.float() + (1.0 if acc else 0.0), 1e-2)
    shard = torch.randn(S, device=d).bfloat16()
    full = torch.zeros(S, device=d, dtype=torch.bfloat16)
    ep += 2
    rc = lib.dsb_symm_all_gather(P(shard), N.ptr(full), ctypes.c_int64(S * 2), P(pads), 0, 1, 1, ctypes.c_uint32(ep), 3, 8, st)
    assert rc == 0, rc
    assert torch.equal(full, shard)
    t = torch.randn(4096, device=d)
    out = torch.empty_like(t)
    ep += 2
    rc = lib.dsb_symm_all_reduce(P(t), N.ptr(out), ctypes.c_int64(t.numel()), N.dt(t), P(pads), 0, 1, 3, ctypes.c_uint32(ep), 4, st)
    assert rc == 0, rc
    close(out, t, 1e-6)
torch.cuda.synchronize()
print(f"sanitize target {which}: ok")
