"""Decode attention over the paged KV cache (Llama-3-8B heads: 32 q / 8 kv, d=128): device-timed, L2 flushed.
DSB200_PAGED_DECODE_MMA=0 selects the CUDA-core kernel, default the tensor-core one."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepspeed_b200.ops.kernels import ragged_ops as R  # noqa: E402

hq, hkv, dd, bs = 32, 8, 128, 128
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
rows = []
for seqs, ctx in ((64, 2048), (8, 8192), (256, 512), (1, 32768)):
    torch.manual_seed(0)
    qkv = torch.randn(seqs, (hq + 2 * hkv) * dd, device="cuda", dtype=torch.bfloat16)
    nb = ctx // bs
    cache = torch.randn(seqs * nb, bs, 2, hkv, dd, device="cuda", dtype=torch.bfloat16)
    bt = torch.arange(seqs * nb, device="cuda", dtype=torch.int32).view(seqs, nb)
    seq_of = torch.arange(seqs, device="cuda", dtype=torch.int32)
    pos_of = torch.full((seqs, ), ctx - 3, device="cuda", dtype=torch.int32)
    out = R.paged_attention(qkv, cache, seq_of, pos_of, bt, hq, hkv, dd, bs)
    # reference on a few (seq, head) pairs
    err = 0.0
    for s_ in (0, seqs - 1):
        kv = cache[bt[s_].long()].reshape(nb * bs, 2, hkv, dd)[:ctx - 2].float()
        q = qkv[s_].view(hq + 2 * hkv, dd)[:hq].float()
        k = kv[:, 0].repeat_interleave(hq // hkv, dim=1)
        v = kv[:, 1].repeat_interleave(hq // hkv, dim=1)
        att = (torch.einsum("hd,nhd->hn", q, k) / dd**0.5).softmax(-1)
        ref = torch.einsum("hn,nhd->hd", att, v).reshape(-1)
        err = max(err, (out[s_].float() - ref).abs().max().item())
    ts = []
    for _ in range(9):
        flush.zero_()
        a, b = torch.cuda.Event(True), torch.cuda.Event(True)
        a.record()
        R.paged_attention(qkv, cache, seq_of, pos_of, bt, hq, hkv, dd, bs)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = sorted(ts)[4]
    kv_bytes = seqs * (ctx - 2) * 2 * hkv * dd * 2
    rows.append({"seqs": seqs, "ctx": ctx, "us": round(ms * 1e3, 1), "kv_GBps": round(kv_bytes / ms / 1e6), "max_abs_err": round(err, 4)})
    print(rows[-1], flush=True)
    del cache
if len(sys.argv) > 1:
    json.dump({"kernel": "mma" if os.environ.get("DSB200_PAGED_DECODE_MMA", "1") != "0" else "cuda-core", "rows": rows},
              open(sys.argv[1], "w"), indent=1)
