"""Micro-benchmark of the biased / block-sparse flash-attention kernels (csrc/cuda/attn_bias.cu) against the PyTorch
formulations they replace: Evoformer attention (SDPA forward + chunked recomputing backward) and block-sparse attention
(gather + SDPA).  CUDA events, 5 warm-up + 20 timed iterations, median.  Writes gpurun_out/attn_bias_bench.json."""
import json
import math
import os
import statistics

import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from _clocks import Clocks


def timeit(fn, warm=5, iters=20):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.median(ts)


def evoformer(B, Nseq, L, H, D, dtype=torch.bfloat16):
    from deepspeed_b200.ops.deepspeed4science import DS4Sci_EvoformerAttention
    from deepspeed_b200.ops.deepspeed4science import evoformer_attn as E
    q, k, v = (torch.randn(B, Nseq, L, H, D, device="cuda", dtype=dtype, requires_grad=True) for _ in range(3))
    mask = torch.zeros(B, Nseq, 1, 1, L, device="cuda", dtype=dtype)
    pair = torch.randn(B, 1, H, L, L, device="cuda", dtype=dtype, requires_grad=True)
    d_o = torch.randn(B, Nseq, L, H, D, device="cuda", dtype=dtype)

    def fwd():
        return DS4Sci_EvoformerAttention(q, k, v, [mask, pair])

    def fwd_bwd():
        out = fwd()
        torch.autograd.grad(out, (q, k, v, pair), d_o)

    rec = {"shape": {"B": B, "N": Nseq, "L": L, "H": H, "D": D}, "dtype": str(dtype)}
    flops_f = 4.0 * B * Nseq * H * L * L * D
    rec["native_fwd_ms"] = timeit(fwd)
    rec["native_fwd_bwd_ms"] = timeit(fwd_bwd)
    orig = E._native_plan
    E._native_plan = lambda *a: None
    try:
        rec["torch_fwd_ms"] = timeit(fwd)
        rec["torch_fwd_bwd_ms"] = timeit(fwd_bwd, warm=2, iters=5)
    finally:
        E._native_plan = orig
    rec["native_fwd_tflops"] = flops_f / rec["native_fwd_ms"] / 1e9
    rec["native_fwd_bwd_tflops"] = 3.5 * flops_f / rec["native_fwd_bwd_ms"] / 1e9
    rec["speedup_fwd"] = rec["torch_fwd_ms"] / rec["native_fwd_ms"]
    rec["speedup_fwd_bwd"] = rec["torch_fwd_bwd_ms"] / rec["native_fwd_bwd_ms"]
    # bytes a perfect kernel must move (Q K V O + pair bias once per batch) vs achieved
    return rec


def sparse(B, H, S, D, block, dtype=torch.bfloat16):
    from deepspeed_b200.ops.sparse_attention import FixedSparsityConfig, SparseSelfAttention
    from deepspeed_b200.ops.sparse_attention import sparse_self_attention as SSA
    cfg = FixedSparsityConfig(num_heads=H, block=block, num_local_blocks=4, num_global_blocks=1)
    attn = SparseSelfAttention(cfg, max_seq_length=S).cuda()
    q, k, v = (torch.randn(B, H, S, D, device="cuda", dtype=dtype, requires_grad=True) for _ in range(3))
    d_o = torch.randn(B, H, S, D, device="cuda", dtype=dtype)
    density = attn.get_layout(S).float().mean().item()

    def fwd():
        return attn(q, k, v)

    def fwd_bwd():
        torch.autograd.grad(fwd(), (q, k, v), d_o)

    rec = {"shape": {"B": B, "H": H, "S": S, "D": D, "block": block}, "layout_density": density}
    rec["native_fwd_ms"] = timeit(fwd)
    rec["native_fwd_bwd_ms"] = timeit(fwd_bwd)
    orig = SSA._native_ok
    SSA._native_ok = lambda *a: False
    try:
        rec["gather_sdpa_fwd_ms"] = timeit(fwd, warm=2, iters=5)
        rec["gather_sdpa_fwd_bwd_ms"] = timeit(fwd_bwd, warm=2, iters=5)
    finally:
        SSA._native_ok = orig
    dense = lambda: torch.nn.functional.scaled_dot_product_attention(q, k, v)
    rec["dense_sdpa_fwd_ms"] = timeit(dense)
    rec["speedup_fwd"] = rec["gather_sdpa_fwd_ms"] / rec["native_fwd_ms"]
    rec["speedup_fwd_bwd"] = rec["gather_sdpa_fwd_bwd_ms"] / rec["native_fwd_bwd_ms"]
    return rec


def main():
    out = {"gpu": torch.cuda.get_device_name(0), "evoformer": [], "sparse": []}
    clk = Clocks()
    for cfg in [(1, 128, 256, 8, 32), (1, 256, 384, 8, 32), (1, 512, 256, 4, 64), (1, 64, 768, 8, 32)]:
        out["evoformer"].append(evoformer(*cfg))
        print(json.dumps(out["evoformer"][-1]))
    for cfg in [(4, 16, 2048, 64, 16), (4, 16, 4096, 64, 64), (2, 16, 8192, 64, 64)]:
        out["sparse"].append(sparse(*cfg))
        print(json.dumps(out["sparse"][-1]))
    out["clocks"] = clk.stop()
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/attn_bias_bench.json", "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
